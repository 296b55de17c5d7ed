# Full single-GPU validation of a round: tests, smoke, bench (cfg5 default + cfg1), reference arm, ncu launch list + full capture.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err; echo "bench cfg5 rc=$?"; tail -3 gpurun_out/bench_cfg5.err
python bench.py --workload cfg1 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; echo "bench cfg1 rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'spmm_csr|bpr_fused' --launch-skip 14 -c 7 -o gpurun_out/prof_cfg5 -f python bench.py --workload cfg5 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_cfg5.ncu-rep --page raw --csv > gpurun_out/prof_cfg5_raw.csv 2>/dev/null
rm -f gpurun_out/prof_cfg5.ncu-rep
python - <<'PY'
import json
for f in ['bench_cfg5','bench_cfg1','bench_ref']:
    try:
        j=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value=%.3e'%j['value'], 'ms/step=%.3f'%j['ms_per_step'], 'e2e=', j.get('e2e') and ('%.3e'%j['e2e']['value']), 'roofline=', j.get('roofline') and (round(j['roofline']['achieved']), round(j['roofline']['frac'],4), round(j['roofline']['avg_launch_us'])), 'clocks=', j.get('clocks') and (j['clocks']['sm_mhz'], j['clocks']['reasons']))
    except Exception as e:
        print(f, 'ERR', e)
PY
