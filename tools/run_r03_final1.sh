# single-GPU validation of the round: full GPU test suite, smoke, the driver's bench invocation, cfg1 bench + launch list, model iterations
mkdir -p gpurun_out/r03
python -m pytest tests -m gpu -x -q > gpurun_out/r03/pytest_gpu_final.log 2>&1; echo "pytest gpu rc=$?"; tail -4 gpurun_out/r03/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r03/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r03/smoke.log
python bench.py > gpurun_out/r03/bench_cfg5_final.json 2> gpurun_out/r03/bench_cfg5_final.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r03/bench_cfg5_final.json').read().strip().splitlines()[-1])
print('cfg5 ms/step %.2f value %.4e e2e ms %.2f frac %.4f clocks %s' % (j['ms_per_step'], j['value'], j['e2e']['ms_per_step'], j['roofline']['frac'], j['clocks']))
s=j.get('secondary'); print('cfg1 ms/step %.4f spmm us %.1f frac %.4f' % (s['ms_per_step'], s['roofline']['avg_launch_us'], s['roofline']['frac']))
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'])
PY
python bench.py --workload cfg1 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r03/bench_cfg1_final.json 2>/dev/null; echo "cfg1 rc=$?"
python tools/bench_models.py > gpurun_out/r03/bench_models.log 2>&1; echo "models rc=$?"; tail -8 gpurun_out/r03/bench_models.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r03/launches_cfg1.csv python bench.py --workload cfg1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-secondary > gpurun_out/r03/ncu_cfg1.log 2>&1; echo "ncu cfg1 rc=$?"
