mkdir -p gpurun_out/r02
python -m pytest tests/test_gpu_spmm_flat.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02/pytest_gpu_h.log
for d in 8 16; do
  python tools/exp_spmm_r02.py cfg5 0 4,5,6 $d > gpurun_out/r02/exp_narrow_na_d$d.log 2>&1; echo "exp d=$d rc=$?"; grep variant gpurun_out/r02/exp_narrow_na_d$d.log
done
