# r02 session 2, call 1: full GPU suite on the rebuilt library + narrow-row kernel A/B (flat v0 / v10 vs r01 v4)
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02/pytest_gpu_b.log
for d in 8 16 32; do
  python tools/exp_spmm_r02.py cfg5 0 4,0,10 $d > gpurun_out/r02/exp_narrow_b_d$d.log 2>&1; echo "exp d=$d rc=$?"; grep variant gpurun_out/r02/exp_narrow_b_d$d.log
done
