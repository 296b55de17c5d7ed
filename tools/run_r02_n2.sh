# 2 GPUs: NCCL parity tests (logs kept) + bench at N=2 with the self-check block and e2e
mkdir -p gpurun_out/r02
python -m pytest tests/test_dist.py -m gpu -q -x > gpurun_out/r02/pytest_dist_n2.log 2>&1; echo "pytest dist rc=$?"; tail -5 gpurun_out/r02/pytest_dist_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02/bench_n2.json 2> gpurun_out/r02/bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/r02/bench_n2.err
python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-secondary --no-e2e > gpurun_out/r02/bench_n1_for_n2.json 2> gpurun_out/r02/bench_n1_for_n2.err; echo "bench n1 rc=$?"
python - <<'PY'
import json
for f in ['bench_n1_for_n2','bench_n2']:
    try:
        j=json.loads(open(f'gpurun_out/r02/{f}.json').read().strip().splitlines()[-1])
        print(f, 'ms/step=%.3f'%j['ms_per_step'], 'check=', j.get('check'), 'e2e=', j.get('e2e') and (round(j['e2e']['ms_per_step'],2), j['e2e']['h2d_bytes_per_step'], j['e2e'].get('loss_last')))
    except Exception as e:
        print(f, 'ERR', e)
PY
