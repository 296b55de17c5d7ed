mkdir -p gpurun_out
N=${1:-8}
python -m pytest tests/test_dist.py -m gpu -q 2>&1 | grep -v "^\s*$" | tail -15
for mode in features rows; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --parallelism $mode > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err; echo "rc=$?"; cat gpurun_out/bench_n${N}_$mode.json; grep -A8 "rank0.*Traceback" gpurun_out/bench_n${N}_$mode.err | head -12
done
