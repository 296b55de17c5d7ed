# N-GPU cfg5 bench sweep: usage  bash tools/run_gpuN.sh N "F:layout F:layout ..."   (F = feature shards, layout = rows|scores)
N=$1; shift
mkdir -p gpurun_out
if [ -n "$RUN_DIST_TESTS" ]; then python -m pytest tests/test_dist.py -m gpu -q -k "$RUN_DIST_TESTS" 2>&1 | grep -v "^\s*$" | tail -8; fi
for spec in $@; do
F=${spec%%:*}; L=${spec##*:}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --feature-shards $F --loss-layout $L > gpurun_out/bench_n${N}_f${F}_$L.json 2> gpurun_out/bench_n${N}_f${F}_$L.err; echo "N=$N F=$F loss=$L rc=$?"
python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_n${N}_f${F}_$L.json') if l.startswith('{')][-1])
    print('  ms/step %.2f  edges/s %.3e  spmm_us %.0f  launches %d  clocks %s' % (j['ms_per_step'], j['value'], j['roofline']['avg_launch_us'], j['gpu_launches'], j['clocks']['sm_mhz']))
except Exception as e:
    print('  ERR', e)
PY
grep -B2 -A12 "Traceback" gpurun_out/bench_n${N}_f${F}_$L.err | head -30
done
