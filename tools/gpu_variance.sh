# Is a slow rank a slow GPU or a busy host?  Every GPU of the box runs the single-GPU bench ALONE (one after the other), then all
# of them run it together under torchrun; the per-GPU step times of both go to gpurun_out/ (tools/gpu_variance.py prints the table).
N=${1:-4}
F="--steps 3 --warmup 3 --no-e2e --no-legs --no-cpu-baseline --no-resample"
for g in $(seq 0 $((N-1))); do
  CUDA_VISIBLE_DEVICES=$g python bench.py $F 2> gpurun_out/solo_$g.err | tail -1 > gpurun_out/solo_$g.json
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N $F 2> gpurun_out/together.err | tail -1 > gpurun_out/together.json
python - <<PY
import json
N=$N
solo=[json.load(open("gpurun_out/solo_%d.json"%g)) for g in range(N)]
tog=json.load(open("gpurun_out/together.json"))
out={"gpus":N,"solo_ms_per_step":[round(s["ms_per_step"],2) for s in solo],"solo_sm_mhz":[s["clocks"]["sm_mhz"] for s in solo],
     "together_ms_per_step":tog["per_rank"]["ms_per_step"],"together_value":tog["value"],"solo_values":[round(s["value"],1) for s in solo],
     "solo_kernels_ms_per_frame":{k:[s["kernels"].get(k,{}).get("ms_per_frame") for s in solo] for k in ("icp","knn_level0","knn_level1","knn_mid","grid_build","radix_sort","ransac_score")}}
json.dump(out,open("gpurun_out/gpu_variance.json","w"),indent=1); print(json.dumps(out))
PY
