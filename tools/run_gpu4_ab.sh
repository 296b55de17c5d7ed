mkdir -p gpurun_out
for mode in overlap serial; do
extra=""; if [ $mode = overlap ]; then extra="--overlap-exchange"; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --loss-layout rows $extra > gpurun_out/bench_n4_rows_$mode.json 2> gpurun_out/bench_n4_rows_$mode.err; echo "$mode rc=$?"; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/bench_n4_rows_$mode.json') if l.startswith('{')][-1])
    print('  ms/step %.2f  edges/s %.3e  spmm_us %.0f  launches %d' % (j['ms_per_step'], j['value'], j['roofline']['avg_launch_us'], j['gpu_launches']))
except Exception as e:
    print('  ERR', e)
PY
grep -B2 -A14 "Traceback" gpurun_out/bench_n4_rows_$mode.err | head -40
done
