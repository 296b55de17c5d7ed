mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -q -x -k "kmeans or ncl_training or row_sparse or c_abi or cfg1_lightgcn_step or cfg2_ncl or cfg3_directau" > gpurun_out/r02/pytest_gpu_e.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02/pytest_gpu_e.log
timeout 600 python tools/bench_models.py > gpurun_out/r02/bench_models_e.log 2>&1; echo "bench_models rc=$?"; tail -12 gpurun_out/r02/bench_models_e.log
