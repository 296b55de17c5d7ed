"""ncu probe: does an L2 access-policy window over the hub item rows change the DRAM traffic of the user-row SpMM half?
Run under `ncu --metrics ...` (development tool)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import exp_narrow as en
from recommendation_b200 import functional as F_, synth
from recommendation_b200.graph import CSRGraph

dev = torch.device("cuda", 0)
U, I, E, d, K = synth.CONFIGS["cfg5"]
users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
n = U + I
a_ui = CSRGraph.from_coo(users.to(torch.int64), items.to(torch.int64) + U, None, U, n, norm="none")
x = torch.randn(n, d, device=dev); y = torch.empty_like(x)
stream = torch.cuda.Stream()
max_persist = en.dev_attr(en.ATTR_MAX_PERSIST)
with torch.cuda.stream(stream):
    for name, persist, win, ratio, miss in (("none", 0, 0, 0.0, 0), ("p79_w79", max_persist, max_persist, 1.0, en.PROP_STREAMING),
                                            ("p79_w32", max_persist, 32 << 20, 1.0, en.PROP_STREAMING),
                                            ("p79_w128_r06", max_persist, 128 << 20, 0.6, en.PROP_NORMAL)):
        if persist:
            en.set_limit(en.LIMIT_PERSIST, persist)
            en.set_window(stream.cuda_stream, x.data_ptr() + U * d * 4, win, ratio, en.PROP_PERSISTING, miss)
        for _ in range(3):   # ncu: the 3rd launch of each group is the warm one
            F_.spmm_raw(a_ui, x, y=y[:U])
        torch.cuda.synchronize()
        print(name, flush=True)
