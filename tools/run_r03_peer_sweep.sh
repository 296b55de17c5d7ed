# CTA budget of the peer movers: per-phase timing at N GPUs for GCF_PEER_CTAS in a few values; usage: bash tools/run_r03_peer_sweep.sh N "64 128 256"
N=${1:-2}; shift
mkdir -p gpurun_out/r03
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "peer or sampler" 2>&1 | tail -2
for C in ${@:-128}; do
GCF_PEER_CTAS=$C timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/phase_dist.py --loss-layout rows --exchange peer > gpurun_out/r03/phase_n${N}_peer_ctas$C.log 2>&1; echo "phase N=$N peer ctas=$C rc=$?"
grep "peer\|step (" gpurun_out/r03/phase_n${N}_peer_ctas$C.log | grep -v "^W" | tail -8
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 tools/peer_bw.py > gpurun_out/r03/peer_bw_n$N.log 2>&1; echo "bw rc=$?"; grep "^G=" gpurun_out/r03/peer_bw_n$N.log
