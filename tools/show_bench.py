import json, sys
f = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/bench.log'
l=[x for x in open(f) if x.startswith('{')]
d=json.loads(l[-1])
print('value', round(d['value'],2), 'ms/step', round(d['ms_per_step'],2), 'e2e', d.get('e2e') and round(d['e2e']['value'],2), 'launches', d['gpu_launches'], d['clocks'])
print('   ', d.get('kernels_note'))
for k,v in d['kernels'].items(): print('    %-22s %8.3f ms/frame share %.3f calls %5d  %8.1f GB/s  frac %.4f'%(k,v['ms_per_frame'],v['share'],v['calls'],v['algorithmic_GBps'],v['frac_of_hbm_peak']))
print(d['frame_stats'])
if d.get('cpu_baseline'): print(d['cpu_baseline'])
