"""Golden fixture for univariate/esrf.py: RUNS THE REFERENCE's own ESRF.buildMotifInducedAdjacencyMatrix,
GraphRecommender.create_joint_sparse_adjaceny, ESRF.Generator and ESRF.Discriminator (esrf.py:955-971,1067-1198) on fixed-seed
inputs, evaluates the loss lines of trainModel (esrf.py:1231-1236,1296-1309) on their outputs, and stores everything as
tests/golden/esrf.npz.

    python tests/golden/make_golden_esrf.py     # needs /root/reference (build container only)

The generator's parameters are made positive before the run: with the reference's N(0, 0.005^2) initialisation
`torch.log(logits + eps)` in gumbel_softmax (esrf.py:1003-1008) sees negative arguments and the whole alternative neighbourhood
is NaN, which would make the comparison vacuous.  The uniform draws of torch.rand_like are logged and replayed by the parity test.
"""
from __future__ import annotations

import sys
import warnings
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
from make_golden import OUT, grads_of, load_ref, t2n  # noqa: E402


def main():
    warnings.filterwarnings("ignore")
    torch.set_num_threads(1)
    rng = np.random.default_rng(61)
    U, I, d, K, nlG, nlD = 120, 80, 16, 4, 2, 2
    u = rng.integers(0, U, 2600)
    i = (rng.zipf(1.25, 2600) - 1) % I
    pairs = sorted(set(zip(u.tolist(), i.tolist())))
    soc = set()
    while len(soc) < 900:
        a, b = rng.choice(U, 2, replace=False)
        soc.add((int(a), int(b)))
    soc = sorted(soc)
    soc += [(b, a) for a, b in soc[:300]]
    soc = sorted(set(soc))
    mod = load_ref("esrf", "univariate/esrf.py")
    S = sp.coo_matrix((np.ones(len(soc), np.float32), ([a for a, _ in soc], [b for _, b in soc])), shape=(U, U), dtype=np.float32)
    Y = sp.coo_matrix((np.ones(len(pairs), np.float32), ([a for a, _ in pairs], [b for _, b in pairs])), shape=(U, I), dtype=np.float32)
    ns = SimpleNamespace(buildSparseRelationMatrix=lambda: S.copy(), buildSparseRatingMatrix=lambda: Y.copy(), num_users=U)
    A = mod.ESRF.buildMotifInducedAdjacencyMatrix(ns).tocsr()
    A.sort_indices()
    data = SimpleNamespace(user={a: a for a in range(U)}, item={b: b for b in range(I)}, trainingData=[(a, b, 1.0) for a, b in pairs])
    joint = mod.GraphRecommender.create_joint_sparse_adjaceny(SimpleNamespace(num_users=U, num_items=I, data=data)).tocsr()
    joint.sort_indices()

    def to_t(M):
        c = M.tocoo()
        return torch.sparse_coo_tensor(np.vstack([c.row, c.col]), c.data.astype(np.float32), c.shape).coalesce()

    torch.manual_seed(13)
    gen = mod.ESRF.Generator(U, d, nlG, K)
    dis = mod.ESRF.Discriminator(U, I, d, nlD)
    with torch.no_grad():
        gen.relation_embeddings.copy_(torch.rand(U, d) * 0.3 + 0.05)
        gen.c_selector.copy_(torch.rand(K, U) * 0.8 + 0.1)
        dis.user_embeddings.mul_(10); dis.item_embeddings.mul_(10)
    draws = []
    real = torch.rand_like

    def logged(x, **kw):
        r = real(x, **kw)
        draws.append(r.clone())
        return r
    seg = 37
    torch.rand_like = logged
    try:
        alt = gen(to_t(A), seg)
    finally:
        torch.rand_like = real
    noise = torch.stack(draws)                                    # [segment rows, K, U]
    user_idx = torch.tensor(rng.integers(0, U, 64)); i_idx = torch.tensor(rng.integers(0, I, 64)); j_idx = torch.tensor(rng.integers(0, I, 64))
    user_idx[:20] = torch.tensor(rng.integers(seg, seg + 100, 20).clip(max=U - 1))   # rows with a generated neighbourhood
    regU, beta = 0.01, 0.2
    # pretraining pass (esrf.py:1224-1236)
    pu, pi = dis(to_t(joint), torch.zeros(U, U), False, 0, K)
    ue, ve, ne = pu[user_idx], pi[i_idx], pi[j_idx]
    pre_pair = -torch.sum(torch.log(torch.sigmoid((ue * ve).sum(1) - (ue * ne).sum(1)) + 1e-10))
    pre_reg = regU * (torch.norm(ue) + torch.norm(ve) + torch.norm(ne))
    g_pre = grads_of(pre_pair + pre_reg, dis.user_embeddings, dis.item_embeddings)
    # adversarial pass (esrf.py:1288-1309)
    su, si = dis(to_t(joint), alt, True, 0, K)
    ue, ve, ne = su[user_idx], si[i_idx], si[j_idx]
    y_ui, y_uj = (ue * ve).sum(1), (ue * ne).sum(1)
    friend = torch.mm(alt[user_idx], su) / K
    y_vi = (friend * ve).sum(1)
    pair = -torch.sum(torch.log(torch.sigmoid(y_ui - y_uj) + 1e-10))
    reg = regU * (torch.norm(ue) + torch.norm(ve) + torch.norm(ne))
    adv = -torch.sum(torch.log(torch.sigmoid(y_ui - y_vi) + 1e-10))
    d_adv = pair + reg + beta * adv
    g_loss = beta * (-torch.sum(torch.log(torch.sigmoid(y_vi - y_ui) + 1e-10)))
    gd = grads_of(d_adv, dis.user_embeddings, dis.item_embeddings)
    gg = grads_of(g_loss, gen.relation_embeddings, gen.c_selector)
    csr = lambda M: dict(indptr=M.indptr.astype(np.int64), indices=M.indices.astype(np.int64), data=M.data.astype(np.float32),
                         shape=np.array(M.shape))
    Sc, Yc = S.tocsr(), Y.tocsr()
    Sc.sort_indices(); Yc.sort_indices()
    np.savez(OUT / "esrf.npz", K=K, n_layers_G=nlG, n_layers_D=nlD, segment=seg, regU=regU, beta=beta,
             **{f"{n}_{k}": v for n, M in (("S", Sc), ("Y", Yc), ("A", A), ("joint", joint)) for k, v in csr(M).items()},
             users=np.array([a for a, _ in pairs]), items=np.array([b for _, b in pairs]),
             gen_relation=t2n(gen.relation_embeddings), gen_projection=t2n(gen.projection_head), gen_selector=t2n(gen.c_selector),
             dis_state_keys=np.array(sorted(dis.state_dict().keys())), gen_state_keys=np.array(sorted(gen.state_dict().keys())),
             dis_user=t2n(dis.user_embeddings), dis_item=t2n(dis.item_embeddings), noise=t2n(noise), alt=t2n(alt),
             user_idx=t2n(user_idx), i_idx=t2n(i_idx), j_idx=t2n(j_idx),
             pre_user=t2n(pu), pre_item=t2n(pi), pre_pair=t2n(pre_pair), pre_reg=t2n(pre_reg), g_pre_user=g_pre[0], g_pre_item=g_pre[1],
             soc_user=t2n(su), soc_item=t2n(si), pair=t2n(pair), reg=t2n(reg), adv=t2n(adv), d_adv=t2n(d_adv), g_loss=t2n(g_loss),
             g_d_user=gd[0], g_d_item=gd[1], g_g_relation=gg[0], g_g_selector=gg[1])
    print("A nnz", A.nnz, "joint nnz", joint.nnz, "alt finite", bool(torch.isfinite(alt).all()), "alt row sums", float(alt[seg].sum()),
          "losses", float(pre_pair), float(pair), float(adv), float(g_loss), "grad finite", bool(np.isfinite(gg[0]).all()))


if __name__ == "__main__":
    main()
