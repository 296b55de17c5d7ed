# Full GPU validation of a round: tests, smoke, bench (cfg5 default + cfg1), reference arm, ncu launch list.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"; cat gpurun_out/bench_cfg5.json; tail -3 gpurun_out/bench_cfg5.err
python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "bench cfg1 rc=$?"; cat gpurun_out/bench_cfg1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
