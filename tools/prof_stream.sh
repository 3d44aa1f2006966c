# `--set full` captures of the streaming kernels of the final build (8 frames per launch, direct launches); raw page exported here
set -x
export KP_PIPE_GRAPH=0 KP_PIPE_BATCH=8 KP_PIPE_SLOTS=1
CMD="python bench.py --steps 1 --warmup 1 --streams 8 --frames-per-step 8 --no-cpu-baseline --no-resample --no-e2e --no-legs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k 'regex:k_unproject|k_rs_scatter|k_rs_hist|k_e_voxel_keys|k_e_voxel_mean|k_eg_mark|k_eg_count|k_eg_scatter|k_bc_scatter|k_e_band_mask' -s 60 -c 40 -o gpurun_out/r02_h_stream $CMD > gpurun_out/ncu5.log 2>&1
ncu -i gpurun_out/r02_h_stream.ncu-rep --page raw --csv > gpurun_out/r02_h_stream.raw.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/
