# N-GPU: parity tests (log kept) + bench with the defaults and with the overlap off; usage: bash tools/run_r03_n2_final.sh N [notest]
N=${1:-2}
mkdir -p gpurun_out/r03
if [ "$2" != "notest" ]; then
timeout 900 python -m pytest tests/test_dist.py -m gpu -q -x > gpurun_out/r03/pytest_dist_n${N}_final.log 2>&1; echo "pytest dist rc=$?"; tail -3 gpurun_out/r03/pytest_dist_n${N}_final.log
fi
for O in auto off; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --overlap-exchange $O > gpurun_out/r03/bench_n${N}_overlap_$O.json 2> gpurun_out/r03/bench_n${N}_overlap_$O.err; echo "bench overlap=$O rc=$?"
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/r03/bench_n${N}_overlap_$O.json') if l.startswith('{')][-1])
    print('  ms/step %.2f  edges/s %.3e  e2e ms %.2f  loss %s' % (j['ms_per_step'], j['value'], j['e2e']['ms_per_step'] if j.get('e2e') else -1, j['check']['loss_after_warmup']))
except Exception as e:
    print('  ERR', e)
PY
done
