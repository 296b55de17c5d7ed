# 2-GPU check: nccl parity tests, then cfg5 bench with the loss on rows / on all-reduced scores.
mkdir -p gpurun_out
python -m pytest tests/test_dist.py tests/test_gpu_parity.py -m gpu -q -k "dist or slices or feature or sharded" 2>&1 | grep -v "^\s*$" | tail -15
for mode in rows scores; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --loss-layout $mode > gpurun_out/bench_n2_loss_$mode.json 2> gpurun_out/bench_n2_loss_$mode.err; echo "rc=$?"; cat gpurun_out/bench_n2_loss_$mode.json | cut -c1-400; grep -B2 -A12 "Traceback" gpurun_out/bench_n2_loss_$mode.err | head -40
done
