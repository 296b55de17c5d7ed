"""Launch one SpMM variant a few times (for ncu).  python tools/prof_spmm.py cfg variant mode[,mode..] [tile]
modes: plain | out | linear | l2norm"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import _lib, functional as F_, synth  # noqa: E402
from recommendation_b200.graph import CSRGraph  # noqa: E402

cfg, variant, modes = sys.argv[1], int(sys.argv[2]), sys.argv[3].split(",")
tile = int(sys.argv[4]) if len(sys.argv) > 4 else 0
dev = torch.device("cuda", 0)
U, I, E, d, K = synth.CONFIGS[cfg]
if len(sys.argv) > 5:
    d = int(sys.argv[5])
if E > 20_000_000:
    users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
else:
    inter = synth.power_law_bipartite(U, I, E, seed=1000 + int(cfg[3:]))
    users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
g = CSRGraph.from_pairs(users, items, U, I, norm="sym", **({"tile_nnz": tile} if tile else {}))
n = U + I
x = torch.randn(n, d, device=dev); a1 = torch.randn(n, d, device=dev); y = torch.empty_like(x)
for _ in range(3):
    for m in modes:
        if m == "plain":
            F_.spmm_raw(g, x, y=y, variant=variant)
        elif m == "out":
            F_.spmm_raw(g, x, out=y, variant=variant)
        elif m == "linear":
            F_.spmm_raw(g, x, out=y, alpha=2.0, post=0.5, addends=[a1], betas=[0.25], variant=variant)
        else:
            F_.spmm_raw(g, x, out=y, epilogue=_lib.EPILOGUE_L2NORM, variant=variant)
torch.cuda.synchronize()
print("ok")
