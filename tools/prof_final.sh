# Round-2 evidence: launch list + `--set full` captures of the top kernels (run under gpurun on ONE GPU).
set -x
export KP_PIPE_GRAPH=0
# 8 frames per launch as in the default bench, one batch slot (ncu serialises the kernels anyway)
export KP_PIPE_BATCH=8 KP_PIPE_SLOTS=1
CMD="python bench.py --steps 1 --warmup 1 --streams 8 --frames-per-step 8 --no-cpu-baseline --no-resample --no-e2e --no-legs"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 390 -c 420 --csv --log-file gpurun_out/r02_g_launches.csv $CMD > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_knn_hist_b -s 6 -c 3 -o gpurun_out/r02_g_knn_hist $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_icp_iter_b -s 64 -c 3 -o gpurun_out/r02_g_icp $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_knn_wbf_b|k_knn_mid_b' -s 10 -c 5 -o gpurun_out/r02_g_knn_wbf $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_unproject|k_rs_scatter|k_e_voxel_keys|k_e_voxel_mean|k_eg_scatter|k_bc_scatter' -s 30 -c 14 -o gpurun_out/r02_g_stream $CMD > gpurun_out/ncu4.log 2>&1
# reports are tens of MB each and gpurun brings back at most 64 MiB: export the pages that are read afterwards, drop the reports
for r in knn_hist icp knn_wbf stream; do ncu -i gpurun_out/r02_g_$r.ncu-rep --page raw --csv > gpurun_out/r02_g_$r.raw.csv 2>/dev/null; done
ncu -i gpurun_out/r02_g_knn_hist.ncu-rep --page source --print-source cuda,sass --csv --kernel-id ::regex:k_knn_hist_b:2 > gpurun_out/r02_g_knn_hist.src.csv 2>/dev/null
ncu -i gpurun_out/r02_g_icp.ncu-rep --page source --print-source cuda,sass --csv --kernel-id ::regex:k_icp_iter_b:2 > gpurun_out/r02_g_icp.src.csv 2>/dev/null
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out/
