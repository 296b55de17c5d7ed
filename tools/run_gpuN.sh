# usage: run_gpuN.sh N "F1 F2 ..." [notest]
mkdir -p gpurun_out
N=$1
if [ "$3" != "notest" ]; then python -m pytest tests/test_dist.py -m gpu -q 2>&1 | grep -v "^\s*$" | tail -15; fi
for F in $2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --feature-shards $F > gpurun_out/bench_n${N}_f$F.json 2> gpurun_out/bench_n${N}_f$F.err; echo "N=$N F=$F rc=$?"; python - <<PY
import json
try:
    j=json.loads(open("gpurun_out/bench_n${N}_f$F.json").read().strip().splitlines()[-1])
    print("  ms/step %.2f  value %.3e  spmm_us %.0f  %s" % (j["ms_per_step"], j["value"], j["roofline"]["avg_launch_us"], j["config"]["parallelism"][:60]))
except Exception as e:
    print("  no json:", e)
PY
grep -A8 "rank0.*Traceback" gpurun_out/bench_n${N}_f$F.err | head -12
done
