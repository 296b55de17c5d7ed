"""Per-phase CUDA-event timing of one feature-sharded LightGCN step (cfg5) on G GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29533 \
        tools/phase_dist.py --loss-layout rows|scores [--steps 5] [--workload cfg5]

Prints, for every phase between two marks of FeatureShardedLightGCNTrainer.step, the mean over the timed steps on rank 0 and
the max over ranks.  Phases are measured on the compute stream: a collective shows up as the time the compute stream WAITS
for it ("exposed"), which is what the step pays.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg5")
    ap.add_argument("--loss-layout", default="rows")
    ap.add_argument("--exchange", default="peer")
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--user-rows", default="owner")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    import bench
    from recommendation_b200.dist import FeatureShardedLightGCNTrainer

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, E, d, K, users, items = bench.make_workload(args.workload, dev)
    tr = FeatureShardedLightGCNTrainer(users, items, U, I, d=d, n_layers=K, lr=0.01, reg_weight=1e-4, seed=1234,
                                       loss_layout=args.loss_layout, exchange=args.exchange, overlap=args.overlap, user_rows=args.user_rows)
    for _ in range(args.warmup):
        tr.step()
    dist.barrier(); torch.cuda.synchronize()
    acc, order, totals = {}, [], []
    for _ in range(args.steps):
        tr.phase_marks = []
        tr.step()
        torch.cuda.synchronize()
        m = tr.phase_marks
        totals.append(m[0][1].elapsed_time(m[-1][1]))
        for (_, a), (label, b) in zip(m[:-1], m[1:]):
            if label not in acc:
                acc[label] = 0.0
                order.append(label)
            acc[label] += a.elapsed_time(b)
    tr.phase_marks = None
    vals = torch.tensor([acc[k] / args.steps for k in order] + [sum(totals) / len(totals)], dtype=torch.float64, device=dev)
    vmax = vals.clone()
    dist.all_reduce(vmax, op=dist.ReduceOp.MAX)
    if rank == 0:
        rec = {"workload": args.workload, "gpus": world, "loss_layout": args.loss_layout, "exchange": tr.exchange, "overlap": tr.overlap, "d_per_rank": tr.dg, "steps": args.steps,
               "phases_ms": [{"phase": k, "rank0": round(float(vals[i]), 3), "max_over_ranks": round(float(vmax[i]), 3)}
                             for i, k in enumerate(order)],
               "step_ms": {"rank0": round(float(vals[-1]), 3), "max_over_ranks": round(float(vmax[-1]), 3)}}
        print(f"== {args.workload}, {world} GPUs, loss on {args.loss_layout} (exchange: {tr.exchange}{', overlapped' if tr.overlap else ''}{', owner-major user rows' if getattr(tr, 'owner_major', False) else ''}), d/G = {tr.dg}")
        for p in rec["phases_ms"]:
            print(f"  {p['phase']:<52s} {p['rank0']:8.3f} ms   (max over ranks {p['max_over_ranks']:8.3f})")
        print(f"  {'step (first to last mark)':<52s} {rec['step_ms']['rank0']:8.3f} ms   (max over ranks {rec['step_ms']['max_over_ranks']:8.3f})")
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(rec) + "\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
