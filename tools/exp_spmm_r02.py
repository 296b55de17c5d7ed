"""r02 SpMM experiments: flat-stream kernel variants x tile sizes against the r01 row-walking kernel (variant 4).

    python tools/exp_spmm_r02.py cfg5 [tile sizes, comma separated] [variants, comma separated]

Every variant is checked against the r01 kernel's output (same arithmetic per row, so the tolerance is tight) for the
three epilogue classes (plain store / linear combination with addends / L2-normalise) before it is timed.
Development tool, CUDA-event timings, inputs >> L2 at cfg5; an L2 flush between launches for the small configs."""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import _lib, functional as F_, synth  # noqa: E402
from recommendation_b200.graph import CSRGraph  # noqa: E402

PEAK = 6534.5


def timeit(fn, iters=7, warm=2, flush=None):
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
    return float(np.median(ts))


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    tiles = [int(t) for t in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
    variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [4, 10, 11, 12, 13, 14]
    dev = torch.device("cuda", 0)
    U, I, E, d, K = synth.CONFIGS[cfg]
    if len(sys.argv) > 4:
        d = int(sys.argv[4])   # row width override (feature-sharded slices)
    if E > 20_000_000:
        users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
    else:
        inter = synth.power_law_bipartite(U, I, E, seed=1000 + int(cfg[3:]))
        users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    n = U + I
    torch.manual_seed(0)
    x = torch.randn(n, d, device=dev)
    a1 = torch.randn(n, d, device=dev)
    flush = None if n * d * 4 > 512 * 2**20 else torch.empty(512 * 2**20, dtype=torch.uint8, device=dev)
    res = {"cfg": cfg}
    ref = {}
    for tile in tiles:
        g = CSRGraph.from_pairs(users, items, U, I, norm="sym", **({"tile_nnz": tile} if tile else {}))
        alg = 8 * g.nnz + 4 * (n + 1) + 8 * n * d
        print(f"== {cfg} tile_nnz={g.tile_nnz}: {g} tiles={g.n_tiles}", flush=True)
        for v in variants:
            y = torch.empty_like(x); o = torch.empty_like(x)
            # plain
            F_.spmm_raw(g, x, y=y, variant=v)
            # linear epilogue: O = 0.5 * (2 * T + 0.25 * a1), Y too
            y2 = torch.empty_like(x)
            F_.spmm_raw(g, x, y=y2, out=o, alpha=2.0, post=0.5, addends=[a1], betas=[0.25], variant=v)
            # full epilogue (L2 normalise)
            o3 = torch.empty_like(x)
            F_.spmm_raw(g, x, out=o3, epilogue=_lib.EPILOGUE_L2NORM, variant=v)
            torch.cuda.synchronize()
            if v == 4 and not ref:
                ref = {"y": y, "o": o, "o3": o3}
                msg = "reference"
            else:
                errs = [float((y - ref["y"]).abs().max()), float((y2 - ref["y"]).abs().max()),
                        float((o - ref["o"]).abs().max()), float((o3 - ref["o3"]).abs().max())]
                scale = float(ref["y"].abs().max())
                ok = all(e_ <= 1e-5 * max(scale, 1.0) for e_ in errs)
                msg = f"max|diff| plain {errs[0]:.2e} linear {errs[1]:.2e}/{errs[2]:.2e} l2norm {errs[3]:.2e} (scale {scale:.2f}) {'OK' if ok else 'MISMATCH'}"
            t_plain = timeit(lambda: F_.spmm_raw(g, x, y=y, variant=v), flush=flush)
            t_lin = timeit(lambda: F_.spmm_raw(g, x, out=o, alpha=2.0, post=0.5, addends=[a1], betas=[0.25], variant=v), flush=flush)
            t_full = timeit(lambda: F_.spmm_raw(g, x, out=o3, epilogue=_lib.EPILOGUE_L2NORM, variant=v), flush=flush)
            o4 = torch.empty_like(x)
            t_out = timeit(lambda: F_.spmm_raw(g, x, out=o4, variant=v), flush=flush)   # same traffic as plain, through the stash path
            print(f"  variant {v:2d}: plain {t_plain*1e3:9.1f} us ({alg/t_plain/1e6:6.0f} GB/s alg = {alg/t_plain/1e6/PEAK*100:5.1f} %)  "
                  f"linear {t_lin*1e3:9.1f} us  l2norm {t_full*1e3:9.1f} us  out-only {t_out*1e3:9.1f} us   {msg}", flush=True)
            res[f"tile{g.tile_nnz}_v{v}"] = {"plain_us": t_plain * 1e3, "linear_us": t_lin * 1e3, "l2norm_us": t_full * 1e3}
        del g
    Path("gpurun_out/r02").mkdir(parents=True, exist_ok=True)
    Path(f"gpurun_out/r02/exp_spmm_{cfg}_d{d}.json").write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
