"""r02: hub-row staging in the SM's L1 (hub flags + L1::evict_last / L1::no_allocate gathers) on the L2-resident configs.

    python tools/exp_hub_r02.py [cfg1,cfg2,cfg3,cfg4] [hub counts, comma separated]

For every config and hub count H: flat kernel with flags at 4 / 3 CTAs per SM (variants 0 / 16) against the same kernel
without flags (variants 10 / 15), plain launch and linear-combination epilogue, with and without an L2 flush between
launches; results must be bit-identical.  CUDA-event timings, median of 9."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import functional as F_, synth  # noqa: E402
from recommendation_b200.graph import CSRGraph  # noqa: E402

PEAK = 6534.5


def timeit(fn, iters=9, warm=3, flush=None):
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
        ts.append(s.elapsed_time(e))
    return float(np.median(ts)) * 1e3


def main():
    cfgs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg1", "cfg2", "cfg3", "cfg4"]
    hubs = [int(h) for h in sys.argv[2].split(",")] if len(sys.argv) > 2 else [256, 640, 1024]
    dev = torch.device("cuda", 0)
    flush = torch.empty(512 * 2**20, dtype=torch.uint8, device=dev)
    res = {}
    for cfg in cfgs:
        U, I, E, d, K = synth.CONFIGS[cfg]
        inter = synth.power_law_bipartite(U, I, E, seed=1000 + int(cfg[3:]))
        users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
        n = U + I
        torch.manual_seed(0)
        x = torch.randn(n, d, device=dev); a1 = torch.randn(n, d, device=dev)
        ref = None
        for H in hubs:
            g = CSRGraph.from_pairs(users, items, U, I, norm="sym", hubs=H)
            alg = 8 * g.nnz + 4 * (n + 1) + 8 * n * d
            share = float((g._hub_col_idx < 0).float().mean()) if g._hub_col_idx is not None else 0.0
            print(f"== {cfg} d={d} nnz={g.nnz} tiles={g.n_tiles} hubs/half={H}: {share * 100:.1f} % of the entries reference a hub column", flush=True)
            for v, name in ((10, "no flags, 4 CTA/SM"), (15, "no flags, 3 CTA/SM"), (0, "hub flags, 4 CTA/SM"), (16, "hub flags, 3 CTA/SM")):
                if H != hubs[0] and v in (10, 15):
                    continue
                y = torch.empty_like(x); o = torch.empty_like(x)
                F_.spmm_raw(g, x, y=y, variant=v)
                F_.spmm_raw(g, x, out=o, alpha=2.0, post=0.5, addends=[a1], betas=[0.25], variant=v)
                torch.cuda.synchronize()
                if ref is None:
                    ref = (y.clone(), o.clone())
                same = torch.equal(y, ref[0]) and torch.equal(o, ref[1])
                t = {}
                for fl_name, fl in (("flushed", flush), ("warm", None)):
                    t[f"plain_{fl_name}"] = timeit(lambda: F_.spmm_raw(g, x, y=y, variant=v), flush=fl)
                    t[f"linear_{fl_name}"] = timeit(lambda: F_.spmm_raw(g, x, out=o, alpha=2.0, post=0.5, addends=[a1], betas=[0.25], variant=v), flush=fl)
                print(f"  {name:22s} plain {t['plain_flushed']:7.1f} us flushed / {t['plain_warm']:7.1f} us warm ({alg / t['plain_warm'] / 1e3 / PEAK * 100:5.1f} % alg)   "
                      f"linear {t['linear_flushed']:7.1f} / {t['linear_warm']:7.1f} us   {'bit-identical' if same else 'MISMATCH'}", flush=True)
                res[f"{cfg}_H{H}_v{v}"] = dict(t, hub_share=share, identical=same)
            del g
    Path("gpurun_out/r02").mkdir(parents=True, exist_ok=True)
    Path("gpurun_out/r02/exp_hub.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
