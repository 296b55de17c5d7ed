# per-phase timing, loss on rows: peer exchange vs NCCL exchange; usage: bash tools/run_r03_phase_px.sh N
N=${1:-2}
mkdir -p gpurun_out/r03
for X in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/phase_dist.py --loss-layout rows --exchange $X --out gpurun_out/r03/phase_n$N.jsonl > gpurun_out/r03/phase_n${N}_rows_$X.log 2>&1; echo "phase N=$N rows/$X rc=$?"
grep -v "^W\|^\[W\|NCCL version\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r03/phase_n${N}_rows_$X.log | tail -24
done
