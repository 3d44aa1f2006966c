set -x
CMD="python bench.py --steps 1 --warmup 1 --streams 1 --no-cpu-baseline --no-resample --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_knn_vbi_b|k_icp_iter_b' -s 20 -c 6 -o gpurun_out/r02_vbi $CMD > gpurun_out/ncu1.log 2>&1
KP_KNN_VBI=0 KP_ICP_VBI=0 $CMD > gpurun_out/plain2.log 2>&1 && \
KP_KNN_VBI=0 KP_ICP_VBI=0 ncu --set full --clock-control none --import-source on -k regex:'k_knn_hist_b|k_icp_iter_b' -s 20 -c 6 -o gpurun_out/r02_grid $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out/*.ncu-rep
