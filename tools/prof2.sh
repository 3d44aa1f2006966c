set -x
export KP_PIPE_GRAPH=0
CMD="python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-resample --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_knn_vbi_b -s 3 -c 3 -o gpurun_out/r02_knn_vbi $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_iter_b -s 33 -c 4 -o gpurun_out/r02_icp_vbi $CMD > gpurun_out/ncu3.log 2>&1
export KP_KNN_VBI=0 KP_ICP_VBI=0
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_knn_hist_b -s 3 -c 3 -o gpurun_out/r02_knn_grid $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_iter_b -s 33 -c 4 -o gpurun_out/r02_icp_grid $CMD > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/r02_*.ncu-rep
