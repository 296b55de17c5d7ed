"""Generate the golden fixtures in this directory by RUNNING THE REFERENCE ITSELF.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

The reference (Cmint22/Recommendation) has no tests, seeds or stored outputs, so its own modules are imported by
path from /root/reference (never copied), fed small fixed-seed inputs, and their inputs / outputs / autograd
gradients are stored as .npz.  tests/test_oracle_golden.py pins oracle/ against these files and the GPU parity
tests replay the same inputs through the CUDA path.  /root/reference does not exist on the GPU box; only the
committed .npz files travel.

Modules used: selfcf.py, ncl.py (with a stub `faiss` module: only run_kmeans needs it), directau.py, ssl4rec.py,
gcl.py.  lightgcn.py cannot be imported (torch_geometric missing) -- see oracle/lightgcn_ref.py.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

REF = Path(os.environ.get("GCF_REFERENCE", "/root/reference"))
OUT = Path(__file__).resolve().parent


def load_ref(name: str, relpath: str, stubs=()):
    for s in stubs:
        if s not in sys.modules:
            m = types.ModuleType(s)
            m.__spec__ = importlib.util.spec_from_loader(s, loader=None)
            sys.modules[s] = m
    spec = importlib.util.spec_from_file_location(f"ref_{name}", REF / relpath)
    mod = importlib.util.module_from_spec(spec)
    cwd = os.getcwd()
    os.chdir("/tmp")  # some reference modules create ./log on use
    try:
        spec.loader.exec_module(mod)
    finally:
        os.chdir(cwd)
    return mod


def tiny_dataset(seed=7, n_users=37, n_items=53, n_inter=420, n_dups=17):
    """[[user_str, item_str, 1.0], ...] with duplicate interactions and zero-padded string ids (so that the
    string sort of ncl.py:60-61 and the first-appearance mapping of selfcf.py:281-288 are both exercised)."""
    rng = np.random.default_rng(seed)
    u = rng.integers(0, n_users, n_inter)
    i = (rng.zipf(1.6, n_inter) - 1) % n_items
    pairs = list(zip(u.tolist(), i.tolist()))
    dup = [pairs[k] for k in rng.integers(0, len(pairs), n_dups)]
    pairs = pairs + dup
    rng.shuffle(pairs)
    train = [[f"u{a:03d}", f"i{b:03d}", 1.0] for a, b in pairs]
    test = [[f"u{a:03d}", f"i{b:03d}", 1.0] for a, b in zip(rng.integers(0, n_users, 40).tolist(), rng.integers(0, n_items, 40).tolist())]
    return train, test


def t2n(t):
    return t.detach().cpu().numpy().copy()   # copy: the array must not alias a parameter that is updated in place later


def grads_of(loss, *params):
    gs = torch.autograd.grad(loss, params, allow_unused=True, retain_graph=True)
    return [t2n(g) if g is not None else np.zeros(tuple(p.shape), np.float32) for g, p in zip(gs, params)]


def main():
    torch.manual_seed(0)
    torch.set_num_threads(1)
    train, test = tiny_dataset()
    d, K = 16, 3

    # ------------------------------------------------------------------ selfcf: sym-normalised graph + encoder
    selfcf = load_ref("selfcf", "selfcf.py")
    data = selfcf.Interaction({}, [list(r) for r in train], [list(r) for r in test])
    U, I = data.user_num, data.item_num
    u_idx = np.array([data.user[r[0]] for r in train], dtype=np.int64)
    i_idx = np.array([data.item[r[1]] for r in train], dtype=np.int64)
    ui = data.ui_adj.tocsr(); ui.sort_indices()
    na = data.norm_adj.tocsr(); na.sort_indices()
    np.savez(OUT / "selfcf_graph.npz", n_users=U, n_items=I, u_idx=u_idx, i_idx=i_idx,
             ui_indptr=ui.indptr, ui_indices=ui.indices, ui_data=ui.data,
             norm_indptr=na.indptr, norm_indices=na.indices, norm_data=na.data.astype(np.float32),
             rowsum=np.asarray(data.ui_adj.sum(1)).ravel().astype(np.float32))
    enc = selfcf.LGCN_Encoder(data, d, K)
    uw, iw = enc.embedding_dict["user_emb"], enc.embedding_dict["item_emb"]
    pu, pi = torch.randn(U, d), torch.randn(I, d)
    ua, ia = enc()
    loss = (ua * pu).sum() + (ia * pi).sum()
    gu, gi = grads_of(loss, uw, iw)
    np.savez(OUT / "selfcf_encoder.npz", n_layers=K, user_w=t2n(uw), item_w=t2n(iw), proj_u=t2n(pu), proj_i=t2n(pi),
             user_all=t2n(ua), item_all=t2n(ia), grad_user_w=gu, grad_item_w=gi)

    # SelfCF_HE forward / loss / history update
    he = selfcf.SelfCF_HE(data, d, 0.3, 2)
    rng = np.random.default_rng(3)
    b_users = rng.integers(0, U, 24).tolist()
    b_items = rng.integers(0, I, 24).tolist()
    his_u0, his_i0 = he.u_target_his.clone(), he.i_target_his.clone()
    out = he({"user": b_users, "item": b_items})
    loss = he.get_loss(out)
    params = [he.online_encoder.embedding_dict["user_emb"], he.online_encoder.embedding_dict["item_emb"],
              he.predictor.weight, he.predictor.bias]
    g = grads_of(loss, *params)
    np.savez(OUT / "selfcf_he.npz", momentum=0.3, n_layers=2, users=np.array(b_users), items=np.array(b_items),
             user_w=t2n(params[0]), item_w=t2n(params[1]), pred_w=t2n(params[2]), pred_b=t2n(params[3]),
             his_u0=t2n(his_u0), his_i0=t2n(his_i0), his_u1=t2n(he.u_target_his), his_i1=t2n(he.i_target_his),
             p_u=t2n(out[0]), t_u=t2n(out[1]), p_i=t2n(out[2]), t_i=t2n(out[3]), loss=t2n(loss),
             g_user_w=g[0], g_item_w=g[1], g_pred_w=g[2], g_pred_b=g[3])

    # ------------------------------------------------------------------ ncl: raw graph + encoder + losses
    ncl = load_ref("ncl", "ncl.py", stubs=("faiss",))
    ndata = ncl.Interaction({}, [tuple(r) for r in train], [tuple(r) for r in test])
    nU, nI = ndata.user_num, ndata.item_num
    raw = ndata.norm_adj  # scipy coo_matrix of ones, duplicates kept, insertion order (ncl.py:76-85)
    n_u_idx = np.array([ndata.user[r[0]] for r in train], dtype=np.int64)
    n_i_idx = np.array([ndata.item[r[1]] for r in train], dtype=np.int64)
    np.savez(OUT / "ncl_graph.npz", n_users=nU, n_items=nI, u_idx=n_u_idx, i_idx=n_i_idx,
             coo_row=raw.row.astype(np.int64), coo_col=raw.col.astype(np.int64), coo_data=raw.data.astype(np.float32))
    nenc = ncl.LGCNEncoder(ndata, d, K)
    nuw, niw = nenc.embedding_dict["user_emb"], nenc.embedding_dict["item_emb"]
    with torch.no_grad():  # raw adjacency grows values by ~deg per layer; keep magnitudes sane
        nuw.mul_(0.5); niw.mul_(0.5)
    ru, ri, all_emb = nenc()
    proj = [torch.randn(nU + nI, d) * (0.1 ** k) for k in range(K + 1)]
    loss = (ru * pu[:nU]).sum() + (ri * pi[:nI]).sum() + sum((e * p).sum() for e, p in zip(all_emb, proj))
    gu, gi = grads_of(loss, nuw, niw)
    np.savez(OUT / "ncl_encoder.npz", n_layers=K, user_w=t2n(nuw), item_w=t2n(niw), proj_u=t2n(pu[:nU]), proj_i=t2n(pi[:nI]),
             proj_layers=np.stack([t2n(p) for p in proj]), user_out=t2n(ru), item_out=t2n(ri),
             all_emb=np.stack([t2n(e) for e in all_emb]), grad_user_w=gu, grad_item_w=gi)

    B = 48
    bu = torch.tensor(rng.integers(0, nU, B)); bp = torch.tensor(rng.integers(0, nI, B)); bn = torch.tensor(rng.integers(0, nI, B))
    ue = torch.randn(B, d, requires_grad=True); pe = torch.randn(B, d, requires_grad=True); ne = torch.randn(B, d, requires_grad=True)
    l_bpr = ncl.bpr_loss(ue, pe, ne)
    g_bpr = grads_of(l_bpr, ue, pe, ne)
    l_reg = ncl.l2_reg_loss(1e-3, ue, pe, ne)
    g_reg = grads_of(l_reg, ue, pe, ne)
    v1 = torch.randn(B, d, requires_grad=True); v2 = torch.randn(B, d, requires_grad=True)
    l_nce = ncl.InfoNCE(v1, v2, 0.2)
    g_nce = grads_of(l_nce, v1, v2)
    l_nce_nocos = ncl.InfoNCE(v1 * 0.3, v2 * 0.3, 0.5, b_cos=False)
    g_nce_nocos = grads_of(l_nce_nocos, v1, v2)
    ctx_e = torch.randn(nU + nI, d, requires_grad=True); ini_e = torch.randn(nU + nI, d, requires_grad=True)
    self_ns = SimpleNamespace(data=SimpleNamespace(user_num=nU, item_num=nI), ssl_temp=0.1, ssl_reg=1e-6, alpha=1.5,
                              proto_reg=8e-8, batch_size=2048)
    l_ssl = ncl.NCLModel.ssl_layer_loss(self_ns, ctx_e, ini_e, bu.tolist(), bp.tolist())
    g_ssl = grads_of(l_ssl, ctx_e, ini_e)
    kc = 5
    self_ns.user_centroids = torch.randn(kc, d); self_ns.item_centroids = torch.randn(kc, d)
    self_ns.user_2cluster = torch.tensor(rng.integers(0, kc, nU)); self_ns.item_2cluster = torch.tensor(rng.integers(0, kc, nI))
    l_proto = ncl.NCLModel.ProtoNCE_loss(self_ns, ini_e, bu.tolist(), bp.tolist())
    g_proto = grads_of(l_proto, ini_e)
    np.savez(OUT / "ncl_losses.npz", n_users=nU, n_items=nI, bu=t2n(bu), bp=t2n(bp), bn=t2n(bn),
             ue=t2n(ue), pe=t2n(pe), ne=t2n(ne), bpr=t2n(l_bpr), g_bpr_u=g_bpr[0], g_bpr_p=g_bpr[1], g_bpr_n=g_bpr[2],
             reg=t2n(l_reg), g_reg_u=g_reg[0], g_reg_p=g_reg[1], g_reg_n=g_reg[2],
             v1=t2n(v1), v2=t2n(v2), nce=t2n(l_nce), g_nce_1=g_nce[0], g_nce_2=g_nce[1],
             nce_nocos=t2n(l_nce_nocos), g_nce_nocos_1=g_nce_nocos[0], g_nce_nocos_2=g_nce_nocos[1],
             ctx=t2n(ctx_e), ini=t2n(ini_e), ssl_temp=0.1, ssl_reg=1e-6, alpha=1.5, ssl=t2n(l_ssl), g_ssl_ctx=g_ssl[0], g_ssl_ini=g_ssl[1],
             proto_reg=8e-8, batch_size=2048, user_centroids=t2n(self_ns.user_centroids), item_centroids=t2n(self_ns.item_centroids),
             user_2cluster=t2n(self_ns.user_2cluster), item_2cluster=t2n(self_ns.item_2cluster), proto=t2n(l_proto), g_proto_ini=g_proto[0])

    # ------------------------------------------------------------------ directau
    dau = load_ref("directau", "directau.py")
    ns = SimpleNamespace(gamma=0.7)
    ns.alignment = lambda x, y: dau.DirectAU.alignment(ns, x, y)
    ns.uniformity = lambda x, t=2: dau.DirectAU.uniformity(ns, x, t)
    xu = torch.randn(B, d, requires_grad=True); xp = torch.randn(B, d, requires_grad=True); xn = torch.randn(B, d, requires_grad=True)
    l_al = ns.alignment(xu, xp); l_un = ns.uniformity(xu)
    l_calc = dau.DirectAU.calculate_loss(ns, xu, xp)
    l_train = dau.DirectAU.calculate_loss(ns, xu, xp) - dau.DirectAU.calculate_loss(ns, xu, xn) + dau.l2_reg_loss(1e-4, xu, xp, xn) / 2048
    np.savez(OUT / "directau_losses.npz", gamma=0.7, xu=t2n(xu), xp=t2n(xp), xn=t2n(xn),
             align=t2n(l_al), g_align_u=grads_of(l_al, xu, xp)[0], g_align_p=grads_of(l_al, xu, xp)[1],
             unif=t2n(l_un), g_unif=grads_of(l_un, xu)[0],
             calc=t2n(l_calc), g_calc_u=grads_of(l_calc, xu, xp)[0], g_calc_p=grads_of(l_calc, xu, xp)[1],
             train=t2n(l_train), g_train_u=grads_of(l_train, xu, xp, xn)[0], g_train_p=grads_of(l_train, xu, xp, xn)[1],
             g_train_n=grads_of(l_train, xu, xp, xn)[2], reg=1e-4, batch_size=2048)

    # ------------------------------------------------------------------ ssl4rec
    ssl = load_ref("ssl4rec", "ssl4rec.py")
    a = torch.randn(B, d, requires_grad=True); b = torch.randn(B, d, requires_grad=True)
    l_bs = ssl.batch_softmax_loss(a, b, 0.2)
    g_bs = grads_of(l_bs, a, b)
    l_n2 = ssl.InfoNCE(a, b, 0.15)
    g_n2 = grads_of(l_n2, a, b)
    sdata = ssl.Interaction([list(r) for r in train], [list(r) for r in test])
    sna = sdata.norm_adj.tocsr(); sna.sort_indices()
    np.savez(OUT / "ssl4rec_losses.npz", a=t2n(a), b=t2n(b), batch_softmax=t2n(l_bs), g_bs_a=g_bs[0], g_bs_b=g_bs[1],
             nce=t2n(l_n2), g_nce_a=g_n2[0], g_nce_b=g_n2[1],
             norm_indptr=sna.indptr, norm_indices=sna.indices, norm_data=sna.data.astype(np.float32),
             u_idx=np.array([sdata.user[r[0]] for r in train]), i_idx=np.array([sdata.item[r[1]] for r in train]),
             n_users=sdata.user_num, n_items=sdata.item_num)

    # ------------------------------------------------------------------ gcl
    gcl = load_ref("gcl", "gcl.py")
    z1 = torch.randn(60, d, requires_grad=True); z2 = torch.randn(60, d, requires_grad=True)
    l_g = gcl.info_nce_loss(z1, z2, 0.2)
    g_g = grads_of(l_g, z1, z2)
    import torch.nn.functional as F
    x = (ue * pe).sum(1) - (ue * ne).sum(1)
    l_gb = -F.logsigmoid(x).mean() + 1e-3 * (ue.norm(2).pow(2) + pe.norm(2).pow(2) + ne.norm(2).pow(2)) / B  # gcl.py:219-223
    g_gb = grads_of(l_gb, ue, pe, ne)
    np.savez(OUT / "gcl_losses.npz", z1=t2n(z1), z2=t2n(z2), info_nce=t2n(l_g), g_z1=g_g[0], g_z2=g_g[1],
             ue=t2n(ue), pe=t2n(pe), ne=t2n(ne), reg_weight=1e-3, bpr_reg=t2n(l_gb), g_u=g_gb[0], g_p=g_gb[1], g_n=g_gb[2])
    # ------------------------------------------------------------------ mhcn (univariate/mhcn.py; `tensorflow` is imported but unused)
    mh = load_ref("mhcn", "univariate/mhcn.py", stubs=("tensorflow",))
    srng = np.random.default_rng(11)
    s_users = sorted({r[0] for r in train})
    social = []
    for _ in range(260):
        a, b = srng.choice(len(s_users), 2, replace=False)
        social.append([s_users[a], s_users[b], 1.0])
    for k in range(0, 120, 2):  # reciprocal edges: the bidirectional motif matrices must not be empty
        social.append([social[k][1], social[k][0], 1.0])
    conf = {"model": {"name": "MHCN"}, "MHCN": {"n_layer": 2, "ss_rate": 0.01}, "emb_size": d, "batch_size": 64, "lr": 0.001,
            "reg_lambda": 1e-4, "max.epoch": 1}
    torch.manual_seed(5)
    m = mh.MHCN(conf, [tuple(r) for r in train], [tuple(r) for r in test], [list(x) for x in social])
    m.build()
    Hs, Hj, Hp = [x.tocsr() for x in m.build_hyper_adj_mats()]
    Rn = mh.Graph.normalize_graph_mat(m.data.interaction_mat).tocsr()
    mu = torch.tensor(srng.integers(0, m.data.user_num, 24)); mv = torch.tensor(srng.integers(0, m.data.item_num, 24))
    mn = torch.tensor(srng.integers(0, m.data.item_num, 24))
    perms = []
    real_randperm = torch.randperm

    def logged_randperm(n, **kw):
        p = real_randperm(n, **kw)
        perms.append(p.clone())
        return p
    torch.randperm = logged_randperm
    try:
        out = m.forward(mu, mv, mn)
    finally:
        torch.randperm = real_randperm
    rec = mh.bpr_loss(out[0], out[1], out[2])
    total = rec + out[3]
    names = sorted(n for n, _ in m.named_parameters())
    params = dict(m.named_parameters())
    gs = grads_of(total, *[params[n] for n in names])
    csr = lambda M: dict(indptr=M.indptr, indices=M.indices, data=M.data.astype(np.float32), shape=np.array(M.shape))
    np.savez(OUT / "mhcn_model.npz", n_layers=2, ss_rate=0.01, user_num=m.data.user_num, item_num=m.data.item_num,
             u_idx=t2n(mu), v_idx=t2n(mv), neg_idx=t2n(mn), perms=np.stack([t2n(p) for p in perms]),
             **{f"Hs_{k}": v for k, v in csr(Hs).items()}, **{f"Hj_{k}": v for k, v in csr(Hj).items()},
             **{f"Hp_{k}": v for k, v in csr(Hp).items()}, **{f"R_{k}": v for k, v in csr(Rn).items()},
             **{f"param__{n}": t2n(params[n]) for n in names}, **{f"grad__{n}": g for n, g in zip(names, gs)},
             batch_user=t2n(out[0]), batch_pos=t2n(out[1]), batch_neg=t2n(out[2]), ss_loss=t2n(out[3]),
             final_user=t2n(out[4]), final_item=t2n(out[5]), rec_loss=t2n(rec))

    # ------------------------------------------------------------------ diffnet (univariate/diffnet.py:1124-1132, forward only needs attributes)
    dn = load_ref("diffnet", "univariate/diffnet.py", stubs=("tensorflow",))
    import scipy.sparse as sp
    nU2, nI2, K2 = m.data.user_num, m.data.item_num, 2
    S_raw = sp.coo_matrix((np.ones(len(social), np.float32), ([m.data.user[x[0]] for x in social], [m.data.user[x[1]] for x in social])),
                          shape=(nU2, nU2)).tocsr()
    S_raw.sum_duplicates()
    rs = np.asarray(S_raw.sum(1)).ravel(); rs[rs == 0] = 1
    S_n = sp.diags(1.0 / rs).dot(S_raw).tocsr().astype(np.float32)          # 1/|followees| weights (diffnet.py:1070-1078)
    A_m = m.data.interaction_mat.tocsr().astype(np.float32)
    to_t = lambda M: torch.sparse_coo_tensor(np.vstack(M.tocoo().coords), M.tocoo().data, M.shape).coalesce()
    ns2 = SimpleNamespace(n_layers=K2, S=to_t(S_n), A=to_t(A_m),
                          user_embeddings=(torch.randn(nU2, d) * 0.05).requires_grad_(True),
                          item_embeddings=(torch.randn(nI2, d) * 0.05).requires_grad_(True),
                          weights=[torch.nn.init.xavier_uniform_(torch.empty(2 * d, d)).requires_grad_(True) for _ in range(K2)])
    fu = dn.DiffNet.forward(ns2)
    du = torch.tensor(srng.integers(0, nU2, 32)); di = torch.tensor(srng.integers(0, nI2, 32)); dj = torch.tensor(srng.integers(0, nI2, 32))
    ue2, ve2, ne2 = fu[du], ns2.item_embeddings[di], ns2.item_embeddings[dj]
    y = (ue2 * ve2).sum(1) - (ue2 * ne2).sum(1)
    dloss = -torch.sum(torch.log(torch.sigmoid(y))) + 1e-4 * (torch.norm(ue2, 2) + torch.norm(ve2, 2) + torch.norm(ne2, 2))  # diffnet.py:1107-1115
    dg = grads_of(dloss, ns2.user_embeddings, ns2.item_embeddings, *ns2.weights)
    np.savez(OUT / "diffnet_model.npz", n_layers=K2, num_users=nU2, num_items=nI2, regU=1e-4,
             **{f"S_{k}": v for k, v in csr(S_n).items()}, **{f"A_{k}": v for k, v in csr(A_m).items()},
             user_w=t2n(ns2.user_embeddings), item_w=t2n(ns2.item_embeddings), weights=np.stack([t2n(w) for w in ns2.weights]),
             u_idx=t2n(du), i_idx=t2n(di), j_idx=t2n(dj), final_user=t2n(fu), loss=t2n(dloss),
             g_user_w=dg[0], g_item_w=dg[1], g_weights=np.stack(dg[2:]))
    # ------------------------------------------------------------------ buir (univariate/buir.py): BUIR_NB without dropout
    bu_mod = load_ref("buir", "univariate/buir.py", stubs=("tensorflow", "faiss"))
    bdata = bu_mod.Interaction({}, [list(r) for r in train], [list(r) for r in test])
    bna = bdata.norm_adj.tocsr(); bna.sort_indices()
    torch.manual_seed(9)
    bm = bu_mod.BUIR_NB(bdata, d, 0.9, 2, 0.2, drop_flag=False)
    with torch.no_grad():  # make online and target differ, as after a few updates
        bm.target_encoder.embedding_dict["user_emb"].mul_(0.7); bm.target_encoder.embedding_dict["item_emb"].add_(0.01)
    brng = np.random.default_rng(13)
    b_users = brng.integers(0, bdata.user_num, 24).tolist(); b_items = brng.integers(0, bdata.item_num, 24).tolist()
    tgt_u0 = t2n(bm.target_encoder.embedding_dict["user_emb"]); tgt_i0 = t2n(bm.target_encoder.embedding_dict["item_emb"])
    bout = bm({"user": b_users, "item": b_items})
    bloss = bm.get_loss(bout)
    bparams = [bm.online_encoder.embedding_dict["user_emb"], bm.online_encoder.embedding_dict["item_emb"], bm.predictor.weight, bm.predictor.bias]
    bg = grads_of(bloss, *bparams)
    bm.update_target(b_users, b_items)
    np.savez(OUT / "buir_nb.npz", n_users=bdata.user_num, n_items=bdata.item_num, momentum=0.9, n_layers=2,
             norm_indptr=bna.indptr, norm_indices=bna.indices, norm_data=bna.data.astype(np.float32),
             users=np.array(b_users), items=np.array(b_items), online_user=t2n(bparams[0]), online_item=t2n(bparams[1]),
             target_user0=tgt_u0, target_item0=tgt_i0, pred_w=t2n(bparams[2]), pred_b=t2n(bparams[3]),
             out_u_online=t2n(bout[0]), out_u_target=t2n(bout[1]), out_i_online=t2n(bout[2]), out_i_target=t2n(bout[3]),
             loss=t2n(bloss), g_user=bg[0], g_item=bg[1], g_pred_w=bg[2], g_pred_b=bg[3],
             target_user1=t2n(bm.target_encoder.embedding_dict["user_emb"]), target_item1=t2n(bm.target_encoder.embedding_dict["item_emb"]))

    # ------------------------------------------------------------------ sept (univariate/sept.py:53-62,220-226)
    sp_mod = load_ref("sept", "univariate/sept.py", stubs=("tensorflow", "faiss"))
    sdat = sp_mod.Interaction({}, [tuple(r) for r in train], [tuple(r) for r in test])
    np.random.seed(31)
    dropped = sp_mod.GraphAugmentor.edge_dropout(sdat.norm_adj, 0.25).tocsr()
    dropped.sort_indices()
    full = sdat.norm_adj.tocsr(); full.sort_indices()
    s_emb = (torch.randn(sdat.user_num + sdat.item_num, d) * 0.3).requires_grad_(True)
    coo = dropped.tocoo()
    adj_t = torch.sparse_coo_tensor(np.vstack([coo.row, coo.col]), coo.data.astype(np.float32), coo.shape).coalesce()
    s_out = sp_mod.SEPT.encoder(SimpleNamespace(n_layers=3), s_emb, adj_t)
    s_proj = torch.randn_like(s_out)
    np.savez(OUT / "sept_encoder.npz", n_layers=3, n_users=sdat.user_num, n_items=sdat.item_num,
             full_indptr=full.indptr, full_indices=full.indices, full_data=full.data.astype(np.float32), full_nnz_nonzero=int(sdat.norm_adj.nnz),
             drop_indptr=dropped.indptr, drop_indices=dropped.indices, drop_data=dropped.data.astype(np.float32),
             emb=t2n(s_emb), out=t2n(s_out), proj=t2n(s_proj), grad=grads_of((s_out * s_proj).sum(), s_emb)[0])

    # ------------------------------------------------------------------ ranking_evaluation (ncl.py:133-178) on random lists
    erng = np.random.default_rng(21)
    eU, eI, eN = 30, 80, 20
    origin, res = {}, {}
    lists = np.stack([erng.permutation(eI)[:eN] for _ in range(eU)])
    test_items = []
    for u in range(eU):
        its = erng.choice(eI, size=int(erng.integers(1, 9)), replace=False)
        origin[f"u{u}"] = {f"i{i}": 1 for i in its}
        res[f"u{u}"] = [(f"i{i}", float(eN - r)) for r, i in enumerate(lists[u])]
        test_items.append(np.array(sorted(its)))
    strings = ncl.ranking_evaluation(origin, res, [5, 10, 20])
    vals = np.array([float(x.split(":")[1]) for x in strings if ":" in x]).reshape(3, 4)
    np.savez(OUT / "eval_metrics.npz", lists=lists, test_ptr=np.concatenate([[0], np.cumsum([len(t) for t in test_items])]),
             test_items=np.concatenate(test_items), top_ns=np.array([5, 10, 20]), measures=vals)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
