mkdir -p gpurun_out
timeout 900 python tools/exp_narrow.py cfg5 gran,rank,window > gpurun_out/exp_narrow.log 2>&1; echo "rc=$?"; tail -80 gpurun_out/exp_narrow.log
