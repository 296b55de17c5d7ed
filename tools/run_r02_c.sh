# r02 session 2, call 2: full GPU suite, default bench line (cfg5 + cfg1 secondary), reference arm
mkdir -p gpurun_out/r02
python -m pytest tests -m gpu -x -q > gpurun_out/r02/pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02/pytest_gpu_c.log
python bench.py > gpurun_out/r02/bench_cfg5_c.json 2> gpurun_out/r02/bench_cfg5_c.err; echo "bench rc=$?"; tail -3 gpurun_out/r02/bench_cfg5_c.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_ref_c.json 2> gpurun_out/r02/bench_ref_c.err; echo "ref rc=$?"
cat gpurun_out/r02/bench_cfg5_c.json gpurun_out/r02/bench_ref_c.json | cut -c1-1500
