"""Motif-induced adjacency build (mhcn.py:340-368) on the GPU vs scipy on the host, at a Douban-Book-like shape (the data
set of the MHCN paper: 2,848 users, 39,586 items, 894,887 interactions, 35,770 social edges) on synthetic power-law data,
and the social-only channels at the cfg4 shape (250,000 users, 1 M social edges, 5 M interactions; Y.Y^T is infeasible there
for either implementation).  Development tool."""
import json, sys, time
from pathlib import Path
import numpy as np, scipy.sparse as sp, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import motif_ref
from recommendation_b200 import motifs, synth
from recommendation_b200.graph import CSRGraph

dev = torch.device("cuda", 0)
res = {}


def social_graph(n, m, seed):
    rng = np.random.default_rng(seed)
    w = np.arange(1, n + 1, dtype=np.float64) ** -0.7
    cdf = np.cumsum(w) / w.sum()
    a = np.minimum(np.searchsorted(cdf, rng.random(2 * m)), n - 1); b = np.minimum(np.searchsorted(cdf, rng.random(2 * m)), n - 1)
    keep = a != b
    key = np.unique(a[keep].astype(np.int64) * n + b[keep])[:m]
    rng.shuffle(key)
    a, b = key // n, key % n
    rec = slice(0, m // 5)                                   # 20 % reciprocated
    a, b = np.concatenate([a, b[rec]]), np.concatenate([b, a[rec]])
    S = sp.coo_matrix((np.ones(len(a), np.float32), (a, b)), shape=(n, n)).tocsr()
    S.data[:] = 1.0
    return S


# ---- Douban-Book-like: full build, GPU vs scipy
U, I, E, M = 2848, 39586, 894887, 35770
inter = synth.power_law_bipartite(U, I, E, seed=77)
Y = sp.coo_matrix((np.ones(E, np.float32), (inter.users, inter.items)), shape=(U, I)).tocsr()
S = social_graph(U, M, 78)
gS, gY = CSRGraph.from_scipy(S, device=dev), CSRGraph.from_scipy(Y, device=dev)
H = motifs.build_hyper_adj_mats(gS, gY); torch.cuda.synchronize()          # warm-up
t0 = time.perf_counter(); H = motifs.build_hyper_adj_mats(gS, gY); torch.cuda.synchronize(); t_gpu = time.perf_counter() - t0
t0 = time.perf_counter(); W = motif_ref.build_hyper_adj_mats(S, Y); t_cpu = time.perf_counter() - t0
same = []
for g, w in zip(H, W):
    w.sort_indices()
    same.append(bool(np.array_equal(g.row_ptr.cpu().numpy(), w.indptr) and np.array_equal(g.col_idx.cpu().numpy(), w.indices)
                     and np.allclose(g.vals.cpu().numpy(), w.data, rtol=3e-7, atol=0)))
res["douban_like"] = {"U": U, "I": I, "E": E, "social": int(S.nnz), "gpu_ms": t_gpu * 1e3, "scipy_ms": t_cpu * 1e3,
                      "nnz": [g.nnz for g in H], "identical_to_scipy": same}
print(json.dumps(res["douban_like"]), flush=True)

# ---- cfg4 shape: the masked products only (social + joint channels)
U, I, E, M = 250_000, 125_000, 5_000_000, 1_000_000
inter = synth.power_law_bipartite(U, I, E, seed=1004)
Y = sp.coo_matrix((np.ones(E, np.float32), (inter.users, inter.items)), shape=(U, I)).tocsr()
S = social_graph(U, M, 79)
gS, gY = CSRGraph.from_scipy(S, device=dev), CSRGraph.from_scipy(Y, device=dev)
St = gS.transpose()
b_vals = gS.vals * motifs.csr_sample(St, gS)
B = motifs.from_coo_sum([motifs.to_coo(gS, b_vals)], U, U, dev); Un = motifs.from_coo_sum([motifs.to_coo(gS, gS.vals - b_vals)], U, U, dev)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
res["cfg4_masked"] = {"social_nnz": gS.nnz, "B_nnz": B.nnz, "U_nnz": Un.nnz,
                      "(U.U)oU^T_ms": timeit(lambda: motifs.masked_product(Un, Un.transpose(), Un.transpose())),
                      "(Y.Y^T)oB_ms": timeit(lambda: motifs.masked_product(gY, gY, B)),
                      "(Y.Y^T)oU_ms": timeit(lambda: motifs.masked_product(gY, gY, Un))}
t0 = time.perf_counter(); Usp = sp.csr_matrix((Un.vals.cpu().numpy(), Un.col_idx.cpu().numpy(), Un.row_ptr.cpu().numpy()), shape=(U, U))
ref = (Usp @ Usp).multiply(Usp.T).tocsr(); res["cfg4_masked"]["scipy_(U.U)oU^T_ms"] = (time.perf_counter() - t0) * 1e3
got = motifs.masked_product(Un, Un.transpose(), Un.transpose())
res["cfg4_masked"]["sum_matches_scipy"] = bool(float(got.sum().item()) == float(ref.sum()))
print(json.dumps(res["cfg4_masked"]), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/bench_motifs.json").write_text(json.dumps(res, indent=1))
