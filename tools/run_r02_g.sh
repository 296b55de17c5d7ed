# profiling call: InfoNCE (post-optimisation ncu), k-means assign, the d = 16 narrow-row SpMM
mkdir -p gpurun_out/r02
python tools/bench_infonce.py > gpurun_out/r02/bench_infonce.log 2>&1; echo "bench_infonce rc=$?"; cat gpurun_out/r02/bench_infonce.log | cut -c1-260
python tools/bench_kmeans.py > gpurun_out/r02/bench_kmeans.log 2>&1; echo "bench_kmeans rc=$?"; cat gpurun_out/r02/bench_kmeans.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lse_stream|grad_stream' --launch-skip 4 -c 4 -o gpurun_out/r02/prof_infonce_bxi -f python tools/bench_infonce.py "B x I" > gpurun_out/r02/ncu_infonce_bxi.log 2>&1; echo "ncu infonce BxI rc=$?"
ncu -i gpurun_out/r02/prof_infonce_bxi.ncu-rep --page raw --csv > gpurun_out/r02/prof_infonce_bxi_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lse_stream|grad_stream' --launch-skip 4 -c 4 -o gpurun_out/r02/prof_infonce_uxu -f python tools/bench_infonce.py "U x U" > gpurun_out/r02/ncu_infonce_uxu.log 2>&1; echo "ncu infonce UxU rc=$?"
ncu -i gpurun_out/r02/prof_infonce_uxu.ncu-rep --page raw --csv > gpurun_out/r02/prof_infonce_uxu_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'km_assign|km_centroid' --launch-skip 30 -c 4 -o gpurun_out/r02/prof_kmeans -f python tools/bench_kmeans.py 1 > gpurun_out/r02/ncu_kmeans.log 2>&1; echo "ncu kmeans rc=$?"
ncu -i gpurun_out/r02/prof_kmeans.ncu-rep --page raw --csv > gpurun_out/r02/prof_kmeans_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr_kernel' --launch-skip 2 -c 1 -o gpurun_out/r02/prof_d16 -f python tools/prof_spmm.py cfg5 0 plain 0 16 > gpurun_out/r02/ncu_d16.log 2>&1; echo "ncu d16 rc=$?"
ncu -i gpurun_out/r02/prof_d16.ncu-rep --page raw --csv > gpurun_out/r02/prof_d16_raw.csv 2>/dev/null
rm -f gpurun_out/r02/prof_infonce_uxu.ncu-rep gpurun_out/r02/prof_d16.ncu-rep
ls -la gpurun_out/r02 | tail -12
