# peer-memory loss exchange at N GPUs: unit tests of the movers, NCCL/peer parity tests, per-phase timing peer vs nccl
N=${1:-2}
mkdir -p gpurun_out/r03
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "peer" > gpurun_out/r03/pytest_peer_unit.log 2>&1; echo "pytest peer unit rc=$?"; tail -3 gpurun_out/r03/pytest_peer_unit.log
timeout 900 python -m pytest tests/test_dist.py -m gpu -q -x -k "feature_sharded or seeded" > gpurun_out/r03/pytest_dist_peer_n$N.log 2>&1; echo "pytest dist rc=$?"; tail -15 gpurun_out/r03/pytest_dist_peer_n$N.log
for X in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/phase_dist.py --loss-layout rows --exchange $X --out gpurun_out/r03/phase_n$N.jsonl > gpurun_out/r03/phase_n${N}_rows_$X.log 2>&1; echo "phase N=$N rows/$X rc=$?"
grep -v "^W\|^\[W\|NCCL version\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r03/phase_n${N}_rows_$X.log | tail -24
done
