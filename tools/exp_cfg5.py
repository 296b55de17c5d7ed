"""cfg5 experiments: L2 residency plan for the SpMM, fused/sorted BPR, whole step.  Development tool."""
import ctypes, json, sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import _lib, functional as F_, synth
from recommendation_b200.graph import CSRGraph
from recommendation_b200.lightgcn import FusedLightGCNTrainer

def timeit(fn, iters=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
dev = torch.device("cuda", 0)
U, I, E, d, K = synth.CONFIGS[cfg]
if E > 20_000_000:
    users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
else:
    inter = synth.power_law_bipartite(U, I, E, seed=1001); users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
lib = _lib.load(); res = {}
g = CSRGraph.from_pairs(users, items, U, I, norm="sym")
n = U + I
x = torch.randn(n, d, device=dev); y = torch.empty_like(x)
alg = 8 * g.nnz + 4 * (n + 1) + 8 * n * d
def report(name, ms):
    res[name] = ms; print(f"{name}: {ms:.3f} ms  ({alg/ms/1e6:.0f} GB/s algorithmic)", flush=True)
report("spmm_plain", timeit(lambda: F_.spmm_raw(g, x, y=y)))
# ---- BPR / step
table = torch.empty(n, d, device=dev); torch.nn.init.xavier_uniform_(table)
for name, kw in (("unsorted_separate", dict(sort_triples=False, fused_bpr=False)), ("sorted_separate", dict(sort_triples=True, fused_bpr=False)),
                 ("unsorted_fused", dict(sort_triples=False, fused_bpr=True)), ("sorted_fused", dict(sort_triples=True, fused_bpr=True))):
    tr = FusedLightGCNTrainer(g, U, I, table.clone(), users, items, n_layers=K, **kw)
    res["step_" + name] = timeit(lambda: tr.step(), iters=4, warm=2)
    print(f"step_{name}: {res['step_'+name]:.2f} ms", flush=True)
    st = _lib.current_stream(); u = U
    if kw["fused_bpr"]:
        fn = lambda: (tr.g_final.zero_(), _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(tr.final[:u]), d, _lib.ptr(tr.final[u:]), d, d, _lib.ptr(tr.pos_u), _lib.ptr(tr.pos_i), _lib.ptr(tr.neg), tr.n_triples, 1, 1, 0.0, 0, 1e-4, 1e-4, 0.0, 1.0, _lib.ptr(tr.loss), None, _lib.ptr(tr.g_final[:u]), d, _lib.ptr(tr.g_final[u:]), d, _lib.ptr(tr.bpr_ws), tr.bpr_ws_bytes, st), "b"))
    else:
        fn = lambda: (_lib.check(lib.gcf_bpr_fwd(_lib.ptr(tr.final[:u]), d, _lib.ptr(tr.final[u:]), d, d, _lib.ptr(tr.pos_u), _lib.ptr(tr.pos_i), _lib.ptr(tr.neg), tr.n_triples, 1, 1, 0.0, 0, 1e-4, 1e-4, 0.0, _lib.ptr(tr.loss), _lib.ptr(tr.coef), _lib.ptr(tr.bpr_ws), tr.bpr_ws_bytes, st), "b"), tr.g_final.zero_(),
                      _lib.check(lib.gcf_bpr_bwd(_lib.ptr(tr.final[:u]), d, _lib.ptr(tr.final[u:]), d, d, _lib.ptr(tr.pos_u), _lib.ptr(tr.pos_i), _lib.ptr(tr.neg), tr.n_triples, 1, _lib.ptr(tr.coef), None, 1e-4, 1e-4, 0.0, _lib.ptr(tr.g_final[:u]), d, _lib.ptr(tr.g_final[u:]), d, st), "b"))
    res["bpr_" + name] = timeit(fn, iters=4, warm=1)
    print(f"bpr_{name} (incl. memset): {res['bpr_'+name]:.2f} ms", flush=True)
    del tr
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/exp_{cfg}.json").write_text(json.dumps(res, indent=1))
