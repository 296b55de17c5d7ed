mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_spmm_flat.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02/pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02/pytest_gpu_f.log
python tools/exp_hub_r02.py > gpurun_out/r02/exp_hub.log 2>&1; echo "exp rc=$?"; cat gpurun_out/r02/exp_hub.log | tail -60
