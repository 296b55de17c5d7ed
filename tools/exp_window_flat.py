"""r03: does an L2 persisting window over the hub rows speed up the flat-stream SpMM, now that it is DRAM-bound?
(r01 measured -24 % DRAM reads for -2 % time on the latency-bound row-walking kernel, profiles/r01_exp_l2_window.md.)
CUDA-event timings of the two halves of the cfg5 operator and of the whole operator, with and without the window."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import exp_narrow as en
from recommendation_b200 import functional as F_, synth
from recommendation_b200.graph import CSRGraph

dev = torch.device("cuda", 0)
U, I, E, d, K = synth.CONFIGS["cfg5"]
if len(sys.argv) > 1:
    d = int(sys.argv[1])          # row width override: 8 / 16 / 32 = the slices of the feature-sharded multi-GPU layouts
sizes = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 32, 64]
users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
n = U + I
full = CSRGraph.from_pairs(users, items, U, I, norm="sym")
cut = int(full.row_ptr[U].item())
def block(r0, r1, e0, e1):
    rp = (full.row_ptr[r0:r1 + 1] - e0).contiguous()
    return CSRGraph(rp, full.col_idx[e0:e1], full.vals[e0:e1], r1 - r0, n, chunk=full.chunk)
a_users, a_items = block(0, U, 0, cut), block(U, n, cut, full.nnz)
x = torch.randn(n, d, device=dev); y = torch.empty_like(x)
stream = torch.cuda.Stream()
max_persist = en.dev_attr(en.ATTR_MAX_PERSIST)
print("max persisting L2:", max_persist >> 20, "MB", flush=True)

def timeit(fn, iters=7, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))

with torch.cuda.stream(stream):
    print(f"d = {d}", flush=True)
    for persist_mb in sizes:
        for miss in ((en.PROP_STREAMING, "streaming"),) if persist_mb else ((0, "-"),):
            row = []
            for name, g, out, base in (("user rows (gather items)", a_users, y[:U], x.data_ptr() + U * d * 4),
                                       ("item rows (gather users)", a_items, y[U:], x.data_ptr())):
                if persist_mb:
                    en.set_limit(en.LIMIT_PERSIST, persist_mb << 20)
                    en.set_window(stream.cuda_stream, base, persist_mb << 20, 1.0, en.PROP_PERSISTING, miss[0])
                else:
                    en.set_limit(en.LIMIT_PERSIST, 0)
                    en.set_window(stream.cuda_stream, base, 0, 0.0, en.PROP_NORMAL, en.PROP_NORMAL)
                t = timeit(lambda: F_.spmm_raw(g, x, y=out))
                row.append(f"{name}: {t:6.3f} ms")
            print(f"persist {persist_mb:3d} MB, miss = {miss[1]:9s}  " + "   ".join(row), flush=True)
    en.set_limit(en.LIMIT_PERSIST, 0)
    en.set_window(stream.cuda_stream, x.data_ptr(), 0, 0.0, en.PROP_NORMAL, en.PROP_NORMAL)
    print(f"whole operator, no window: {timeit(lambda: F_.spmm_raw(full, x, y=y)):6.3f} ms")
