mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -q > gpurun_out/r02/pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02/pytest_gpu_d.log
