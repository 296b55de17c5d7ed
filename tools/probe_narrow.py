"""ncu probe of the narrow-row SpMM (d = 8: 32 B rows, one rank of the 8-way feature-sharded layout) on the cfg5 graph."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import functional as F_, synth
from recommendation_b200.graph import CSRGraph

dev = torch.device("cuda", 0)
U, I, E, d, K = synth.CONFIGS["cfg5"]
users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
g = CSRGraph.from_pairs(users, items, U, I, norm="sym")
x = torch.randn(U + I, 8, device=dev); y = torch.empty_like(x)
for _ in range(3):
    F_.spmm_raw(g, x, y=y)
torch.cuda.synchronize()
