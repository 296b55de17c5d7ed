# 8-GPU record: NCCL/peer parity tests (log kept), per-phase timing with and without overlap, bench lines
N=${1:-8}
mkdir -p gpurun_out/r03
timeout 600 python -m pytest tests/test_dist.py -m gpu -q -x -k "feature_sharded or (seeded and rows)" > gpurun_out/r03/pytest_dist_n$N.log 2>&1; echo "pytest dist rc=$?"; tail -4 gpurun_out/r03/pytest_dist_n$N.log
for O in "" "--overlap"; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/phase_dist.py --loss-layout rows --exchange peer $O --out gpurun_out/r03/phase_n$N.jsonl > gpurun_out/r03/phase_n${N}_peer_final$O.log 2>&1; echo "phase $O rc=$?"
grep -v "^W\|^\[W\|NCCL version\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r03/phase_n${N}_peer_final$O.log | tail -16
done
for O in "" "--overlap-exchange"; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 $O > gpurun_out/r03/bench_n${N}$O.json 2> gpurun_out/r03/bench_n${N}$O.err; echo "bench $O rc=$?"
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/r03/bench_n${N}$O.json') if l.startswith('{')][-1])
    print('  ms/step %.2f  edges/s %.3e  e2e ms %.2f  check %s' % (j['ms_per_step'], j['value'], j['e2e']['ms_per_step'] if j.get('e2e') else -1, j.get('check')))
except Exception as e:
    print('  ERR', e)
PY
done
