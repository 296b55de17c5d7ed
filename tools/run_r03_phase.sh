# per-phase timing of the feature-sharded step at N GPUs (both loss layouts); usage: bash tools/run_r03_phase.sh N
N=${1:-2}
mkdir -p gpurun_out/r03
for L in rows scores; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/phase_dist.py --loss-layout $L --out gpurun_out/r03/phase_n$N.jsonl > gpurun_out/r03/phase_n${N}_$L.log 2>&1; echo "phase N=$N $L rc=$?"
grep -v "^W\|^\[W\|NCCL version" gpurun_out/r03/phase_n${N}_$L.log | tail -22
done
