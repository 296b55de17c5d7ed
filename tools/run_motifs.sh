mkdir -p gpurun_out
python -m pytest tests/test_gpu_models.py -m gpu -q -x -k "sparse_product or hyper_adj" > gpurun_out/pytest_motifs.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/pytest_motifs.log
