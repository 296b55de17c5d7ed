mkdir -p gpurun_out
python tools/bench_infonce.py 2>&1 | tee gpurun_out/bench_infonce.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'lse_stream|grad_stream' --launch-skip 2 -c 3 -o gpurun_out/prof_infonce -f python tools/bench_infonce.py "B x I" > gpurun_out/ncu_infonce.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_infonce.ncu-rep --page raw --csv > gpurun_out/prof_infonce_raw.csv 2>/dev/null
ls -la gpurun_out | grep infonce
