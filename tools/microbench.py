"""Kernel micro-benchmarks on one GPU (CUDA events, warm + cold L2).  Development tool, not the contract bench."""
from __future__ import annotations

import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import _lib, functional as F_, synth  # noqa: E402
from recommendation_b200.graph import CSRGraph  # noqa: E402
from recommendation_b200.lightgcn import FusedLightGCNTrainer  # noqa: E402

PEAK = 6534.5  # GB/s, MEASURED_PEAKS.json


def timeit(fn, iters=20, warm=3, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts = np.array(ts)
    return float(np.median(ts)), float(ts.min())


def spmm_bytes(n, nnz, d):
    return 8 * nnz + 4 * (n + 1) + 8 * n * d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="cfg1")
    ap.add_argument("--chunks", default="256")
    ap.add_argument("--skip-step", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    U, I, E, d, K = synth.CONFIGS[args.cfg]
    t0 = time.time()
    if E > 20_000_000:
        users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
    else:
        inter = synth.power_law_bipartite(U, I, E, seed=1001)
        users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    torch.cuda.synchronize(); print(f"graph gen {time.time()-t0:.2f}s", flush=True)
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)
    res = {"cfg": args.cfg}
    for chunk in [int(c) for c in args.chunks.split(",")]:
        t0 = time.time()
        g = CSRGraph.from_pairs(users, items, U, I, norm="sym", chunk=chunk)
        torch.cuda.synchronize(); tb = time.time() - t0
        print(f"chunk={chunk}: {g} build {tb*1e3:.1f} ms", flush=True)
        n = U + I
        x = torch.randn(n, d, device=dev); y = torch.empty_like(x)
        B = spmm_bytes(n, g.nnz, d)
        for variant in (0, 1, 2, 3):
            fn = lambda: F_.spmm_raw(g, x, y=y, variant=variant)
            warm, wmin = timeit(fn)
            cold, cmin = timeit(fn, flush=flush)
            print(f"  spmm d={d} variant={variant}: warm {warm:.1f} us ({B/warm/1e3:.0f} GB/s alg, {B/warm/1e3/PEAK*100:.1f}% of measured HBM) "
                  f"cold {cold:.1f} us ({B/cold/1e3:.0f} GB/s, {B/cold/1e3/PEAK*100:.1f}%) gather-model {(g.nnz*d*4+g.nnz*8+n*d*4)/warm/1e3:.0f} GB/s", flush=True)
            res[f"spmm_c{chunk}_v{variant}"] = {"warm_us": warm, "cold_us": cold}
    if not args.skip_step:
        g = CSRGraph.from_pairs(users, items, U, I, norm="sym")
        table = torch.empty(U + I, d, device=dev); torch.nn.init.xavier_uniform_(table)
        tr = FusedLightGCNTrainer(g, U, I, table, users, items, n_layers=K)
        lib = _lib.load(); st = _lib.current_stream()
        def part(name, fn):
            warm, _ = timeit(fn, iters=10)
            cold, _ = timeit(fn, iters=10, flush=flush)
            print(f"  {name}: warm {warm:.1f} us cold {cold:.1f} us", flush=True)
            res[name] = {"warm_us": warm, "cold_us": cold}
        part("step", lambda: tr.step())
        part("prop_fwd", lambda: _lib.check(lib.gcf_propagate_fwd(g.struct_ref(), d, K, _lib.ptr(tr.table), _lib.ptr_array(tr.layers), _lib.ptr(tr.final), 1.0, _lib.ptr(tr.ws), tr.ws_bytes, st), "f"))
        u = U
        part("bpr_fwd", lambda: _lib.check(lib.gcf_bpr_fwd(_lib.ptr(tr.final[:u]), d, _lib.ptr(tr.final[u:]), d, d, _lib.ptr(tr.pos_u), _lib.ptr(tr.pos_i), _lib.ptr(tr.neg), tr.n_triples, 1, 1, 0.0, 0, 1e-4, 1e-4, 0.0, _lib.ptr(tr.loss), _lib.ptr(tr.coef), _lib.ptr(tr.bpr_ws), tr.bpr_ws_bytes, st), "b"))
        part("bpr_bwd", lambda: _lib.check(lib.gcf_bpr_bwd(_lib.ptr(tr.final[:u]), d, _lib.ptr(tr.final[u:]), d, d, _lib.ptr(tr.pos_u), _lib.ptr(tr.pos_i), _lib.ptr(tr.neg), tr.n_triples, 1, _lib.ptr(tr.coef), None, 1e-4, 1e-4, 0.0, _lib.ptr(tr.g_final[:u]), d, _lib.ptr(tr.g_final[u:]), d, st), "b"))
        part("prop_bwd", lambda: _lib.check(lib.gcf_propagate_bwd(g.struct_ref(), d, K, _lib.ptr(tr.g_final), None, 1.0, _lib.ptr(tr.ping), _lib.ptr(tr.pong), _lib.ptr(tr.g_x0), _lib.ptr(tr.ws), tr.ws_bytes, st), "p"))
        part("adam", lambda: _lib.check(lib.gcf_adam_step(_lib.ptr(tr.table), _lib.ptr(tr.g_x0), _lib.ptr(tr.exp_avg), _lib.ptr(tr.exp_avg_sq), tr.table.numel(), 0.01, 0.9, 0.999, 1e-8, 0.0, 0, 5, st), "a"))
        part("sampler", lambda: _lib.check(lib.gcf_sample_negatives(1, 2, None, tr.n_triples, 1, I, None, None, 1, _lib.ptr(tr.neg), st), "s"))
        part("memset", lambda: tr.g_final.zero_())
        # sorted-by-user triples: how much does run aggregation buy?
        order = torch.argsort(users * I + items)
        tr2 = FusedLightGCNTrainer(g, U, I, table.clone(), users[order], items[order], n_layers=K)
        part("step_sorted", lambda: tr2.step())
        res["edges_per_s_warm"] = E / (res["step"]["warm_us"] * 1e-6)
        res["edges_per_s_sorted_warm"] = E / (res["step_sorted"]["warm_us"] * 1e-6)
    Path("gpurun_out").mkdir(exist_ok=True)
    Path(f"gpurun_out/microbench_{args.cfg}.json").write_text(json.dumps(res, indent=1))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
