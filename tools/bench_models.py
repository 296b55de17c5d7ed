"""Mini-batch model steps at the BASELINE cfg2 / cfg3 shapes (NCL, DirectAU, SelfCF): ms per training iteration through the
drop-in classes.  Development tool; the contract bench is bench.py."""
import json, sys, time
from pathlib import Path
from types import SimpleNamespace
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import encoders, functional as F_, losses, ncl as ncl_mod, synth, sampling
from recommendation_b200.graph import CSRGraph

dev = torch.device("cuda", 0)


def data_for(cfg, raw):
    inter, d, K = synth.config_graph(cfg)
    U, I = inter.n_users, inter.n_items
    u, i = inter.users, inter.items
    adj = sp.coo_matrix((np.ones(2 * len(u), np.float32), (np.concatenate([u, i + U]), np.concatenate([i + U, u]))), shape=(U + I, U + I))
    data = SimpleNamespace(user_num=U, item_num=I, norm_adj=adj)
    if not raw:  # selfcf / ssl4rec normalise; ncl / directau keep the raw adjacency
        data._gcf_graph = CSRGraph.from_scipy(adj, norm="sym", device=dev)
    smp = sampling.PairwiseSampler(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), U, I)
    return data, smp, d, K


def timed(step, batches, warm=3, iters=20):
    ts = []
    for n, b in enumerate(batches):
        if n == warm: torch.cuda.synchronize(); t0 = time.perf_counter()
        step(b)
        if n == warm + iters - 1: break
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / iters * 1e3


res = {}
data, smp, d, K = data_for("cfg2", raw=True)
model = encoders.LGCNEncoder(data, d, K)
with torch.no_grad():
    for p in model.parameters(): p.mul_(1e-3)   # raw adjacency: keep activations finite
ncl = losses.NCLLosses(data.user_num, data.item_num, 0.1, 1e-6, 1.5, 8e-8, 4096)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
t0 = time.perf_counter(); k = ncl_mod.e_step(model, ncl, 1000); torch.cuda.synchronize()
res["ncl_e_step_ms (k-means of 52,643 + 91,599 rows, k=%d)" % k] = (time.perf_counter() - t0) * 1e3
torch.cuda.synchronize(); t0 = time.perf_counter(); ncl_mod.e_step(model, ncl, 1000); torch.cuda.synchronize()
res["ncl_e_step_ms, warm (second call)"] = (time.perf_counter() - t0) * 1e3
res["ncl_step_ms (cfg2, B=4096, 3 layers, BPR + ssl_layer + ProtoNCE + Adam)"] = timed(
    lambda b: ncl_mod.ncl_step(model, ncl, opt, b, 1e-4, 4096, 1), smp.batches(4096))
# the reference's own iteration: e_step() after every batch (ncl.py:324)
res["ncl_step_ms with the E-step inside every batch (ncl.py:324: refresh_clusters=True)"] = timed(
    lambda b: ncl_mod.ncl_step(model, ncl, opt, b, 1e-4, 4096, 1, k=k, refresh_clusters=True), smp.batches(4096), warm=2, iters=10)
del model, opt
data, smp, d, K = data_for("cfg3", raw=True)
model = encoders.LGCNEncoder(data, d, K)
with torch.no_grad():
    for p in model.parameters(): p.mul_(1e-3)
dau = losses.DirectAULosses(0.7)
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
def dau_step(b):
    u, p_, n_ = b
    ue, ie, _ = model()
    a, c, e = F_.gather_rows(ue, u), F_.gather_rows(ie, p_), F_.gather_rows(ie, n_)
    loss = dau.calculate_loss(a, c) - dau.calculate_loss(a, e) + losses.l2_reg_loss(1e-4, a, c, e) / 2048
    opt.zero_grad(); loss.backward(); opt.step()
res["directau_step_ms (cfg3, B=2048, d=128, 2 layers)"] = timed(dau_step, smp.batches(2048))
del model, opt
data, smp, d, K = data_for("cfg3", raw=False)
he = encoders.SelfCF_HE(data, d, 0.05, K)
opt = torch.optim.Adam(he.parameters(), lr=1e-3)
def he_step(b):
    out = he({"user": b[0], "item": b[1]})
    loss = he.get_loss(out)
    opt.zero_grad(); loss.backward(); opt.step()
res["selfcf_he_step_ms (cfg3, B=1024, d=128, 2 layers)"] = timed(he_step, smp.batches(1024))
print(json.dumps(res, indent=1))
