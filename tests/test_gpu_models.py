"""GPU parity of the drop-in model classes against fixtures produced by the reference's own classes
(tests/golden/make_golden.py): same parameters in, same outputs / losses / gradients out."""
from types import SimpleNamespace

import numpy as np
import pytest
import scipy.sparse as sp
import torch

from recommendation_b200 import _lib as _lib_mod, encoders, losses, sampling
from recommendation_b200.graph import CSRGraph

pytestmark = pytest.mark.gpu
RTOL = 1e-3


def _close(got, want, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(got.detach().cpu().numpy(), want, rtol=rtol, atol=atol)


def _selfcf_data(g):
    U, I = int(g["n_users"]), int(g["n_items"])
    norm = sp.csr_matrix((g["norm_data"], g["norm_indices"], g["norm_indptr"]), shape=(U + I, U + I))
    return SimpleNamespace(user_num=U, item_num=I, norm_adj=norm)


def _ncl_data(g):
    U, I = int(g["n_users"]), int(g["n_items"])
    raw = sp.coo_matrix((g["coo_data"], (g["coo_row"], g["coo_col"])), shape=(U + I, U + I))  # duplicates kept (ncl.py:76-85)
    return SimpleNamespace(user_num=U, item_num=I, norm_adj=raw)


def _load(enc, z, cuda):
    sd = {"embedding_dict.user_emb": torch.from_numpy(z["user_w"]), "embedding_dict.item_emb": torch.from_numpy(z["item_w"])}
    assert sorted(enc.state_dict().keys()) == sorted(sd.keys())   # state_dict keys are part of the interface
    enc.load_state_dict(sd)
    return enc.embedding_dict["user_emb"], enc.embedding_dict["item_emb"]


def test_lgcn_encoder_selfcf_fixture(cuda, golden):
    z, data = golden("selfcf_encoder"), _selfcf_data(golden("selfcf_graph"))
    enc = encoders.LGCN_Encoder(data, z["user_w"].shape[1], int(z["n_layers"]))
    uw, iw = _load(enc, z, cuda)
    ua, ia = enc()
    _close(ua, z["user_all"]); _close(ia, z["item_all"])
    ((ua * torch.from_numpy(z["proj_u"]).to(cuda)).sum() + (ia * torch.from_numpy(z["proj_i"]).to(cuda)).sum()).backward()
    _close(uw.grad, z["grad_user_w"]); _close(iw.grad, z["grad_item_w"])


def test_lgcn_encoder_ncl_fixture(cuda, golden):
    z, data = golden("ncl_encoder"), _ncl_data(golden("ncl_graph"))
    enc = encoders.LGCNEncoder(data, z["user_w"].shape[1], int(z["n_layers"]))
    uw, iw = _load(enc, z, cuda)
    ru, ri, all_emb = enc()
    assert len(all_emb) == int(z["n_layers"]) + 1
    _close(ru, z["user_out"], atol=1e-5); _close(ri, z["item_out"], atol=1e-5)
    for k, e in enumerate(all_emb):
        _close(e, z["all_emb"][k], atol=1e-5 * max(1.0, np.abs(z["all_emb"][k]).max()))
    proj = torch.from_numpy(z["proj_layers"]).to(cuda)
    loss = (ru * torch.from_numpy(z["proj_u"]).to(cuda)).sum() + (ri * torch.from_numpy(z["proj_i"]).to(cuda)).sum()
    loss = loss + sum((e * proj[k]).sum() for k, e in enumerate(all_emb))
    loss.backward()
    _close(uw.grad, z["grad_user_w"], atol=1e-4); _close(iw.grad, z["grad_item_w"], atol=1e-4)


def test_selfcf_he_fixture(cuda, golden):
    z, data = golden("selfcf_he"), _selfcf_data(golden("selfcf_graph"))
    he = encoders.SelfCF_HE(data, z["user_w"].shape[1], float(z["momentum"]), int(z["n_layers"]))
    want_keys = ["online_encoder.embedding_dict.item_emb", "online_encoder.embedding_dict.user_emb", "predictor.bias", "predictor.weight"]
    assert sorted(he.state_dict().keys()) == want_keys          # history buffers are NOT in the state_dict
    he.load_state_dict({"online_encoder.embedding_dict.user_emb": torch.from_numpy(z["user_w"]),
                        "online_encoder.embedding_dict.item_emb": torch.from_numpy(z["item_w"]),
                        "predictor.weight": torch.from_numpy(z["pred_w"]), "predictor.bias": torch.from_numpy(z["pred_b"])})
    he.u_target_his = torch.from_numpy(z["his_u0"]).to(cuda); he.i_target_his = torch.from_numpy(z["his_i0"]).to(cuda)
    out = he({"user": z["users"].tolist(), "item": z["items"].tolist()})      # Python lists, as the reference passes
    for got, key in zip(out, ("p_u", "t_u", "p_i", "t_i")):
        _close(got, z[key], atol=1e-5)
    loss = he.get_loss(out)
    _close(loss, z["loss"], rtol=1e-4)
    loss.backward()
    enc = he.online_encoder
    _close(enc.embedding_dict["user_emb"].grad, z["g_user_w"], atol=1e-6); _close(enc.embedding_dict["item_emb"].grad, z["g_item_w"], atol=1e-6)
    _close(he.predictor.weight.grad, z["g_pred_w"], atol=1e-6); _close(he.predictor.bias.grad, z["g_pred_b"], atol=1e-6)
    _close(he.u_target_his, z["his_u1"], atol=1e-5); _close(he.i_target_his, z["his_i1"], atol=1e-5)
    pu, uo, pi, io = he.get_embedding()
    assert pu.shape == uo.shape == (data.user_num, z["user_w"].shape[1]) and not pu.requires_grad


def test_reference_named_losses(cuda, golden):
    z = golden("ncl_losses")
    T = lambda a: torch.from_numpy(np.asarray(a)).to(cuda)
    L = lambda a: T(a).float().requires_grad_(True)
    ue, pe, ne = L(z["ue"]), L(z["pe"]), L(z["ne"])
    loss = losses.bpr_loss(ue, pe, ne)
    _close(loss, z["bpr"], rtol=1e-4)
    loss.backward()
    _close(ue.grad, z["g_bpr_u"], atol=1e-7); _close(pe.grad, z["g_bpr_p"], atol=1e-7); _close(ne.grad, z["g_bpr_n"], atol=1e-7)
    ue, pe, ne = L(z["ue"]), L(z["pe"]), L(z["ne"])
    reg = losses.l2_reg_loss(1e-3, ue, pe, ne)
    _close(reg, z["reg"], rtol=1e-5)
    ncl = losses.NCLLosses(int(z["n_users"]), int(z["n_items"]), float(z["ssl_temp"]), float(z["ssl_reg"]), float(z["alpha"]),
                           float(z["proto_reg"]), int(z["batch_size"]))
    ctx, ini = L(z["ctx"]), L(z["ini"])
    l = ncl.ssl_layer_loss(ctx, ini, z["bu"].tolist(), z["bp"].tolist())
    np.testing.assert_allclose(l.item(), z["ssl"], rtol=2e-2)
    ncl.user_centroids, ncl.item_centroids = T(z["user_centroids"]), T(z["item_centroids"])
    ncl.user_2cluster, ncl.item_2cluster = T(z["user_2cluster"]), T(z["item_2cluster"])
    ini2 = L(z["ini"])
    l = ncl.ProtoNCE_loss(ini2, z["bu"].tolist(), z["bp"].tolist())
    np.testing.assert_allclose(l.item(), z["proto"], rtol=2e-2)
    l.backward()
    want = z["g_proto_ini"]
    np.testing.assert_allclose(ini2.grad.cpu().numpy(), want, rtol=2e-2, atol=2e-2 * np.abs(want).max())
    zd = golden("directau_losses")
    dau = losses.DirectAULosses(float(zd["gamma"]))
    xu, xp, xn = L(zd["xu"]), L(zd["xp"]), L(zd["xn"])
    train = dau.calculate_loss(xu, xp) - dau.calculate_loss(xu, xn) + losses.l2_reg_loss(float(zd["reg"]), xu, xp, xn) / int(zd["batch_size"])
    np.testing.assert_allclose(train.item(), zd["train"], rtol=2e-2, atol=2e-3)
    train.backward()
    for t, k in ((xu, "g_train_u"), (xp, "g_train_p"), (xn, "g_train_n")):
        np.testing.assert_allclose(t.grad.cpu().numpy(), zd[k], rtol=2e-2, atol=2e-2 * np.abs(zd[k]).max())
    np.testing.assert_allclose(dau.alignment(xu, xp).item(), zd["align"], rtol=1e-4)
    np.testing.assert_allclose(dau.uniformity(xu).item(), zd["unif"], rtol=2e-2)


def test_dnn_encoder_and_grace_interfaces(cuda):
    data = SimpleNamespace(user_num=50, item_num=70)
    torch.manual_seed(0)
    m = encoders.DNNEncoder(data, 32, 0.1, 0.2, 2)
    keys = sorted(m.state_dict().keys())
    assert keys == sorted(["initial_user", "initial_item", "user_net.0.weight", "user_net.0.bias", "user_net.2.weight", "user_net.2.bias",
                           "item_net.0.weight", "item_net.0.bias", "item_net.2.weight", "item_net.2.bias"])
    u, i = [1, 2, 3, 3], [5, 6, 7, 7]
    a, b = m(u, i)
    assert a.shape == b.shape == (4, 128)
    m.eval()
    cl = m.cal_cl_loss(i)           # eval mode: both dropout views are identical -> InfoNCE of a batch with itself
    emb = m.item_net(m.initial_item[torch.tensor(i, device=cuda)])
    en = torch.nn.functional.normalize(emb.double(), dim=1)
    want = -torch.diag(torch.log_softmax(en @ en.T / 0.2, dim=1)).mean()
    np.testing.assert_allclose(cl.item(), want.item(), rtol=2e-2)
    (cl + losses.batch_softmax_loss(a, b, 0.2)).backward()
    assert m.initial_item.grad is not None and m.initial_user.grad is not None
    g = encoders.GRACEModel(50, 70, emb_size=32, num_layers=2, proj_dim=16).to(cuda)
    assert "convs.1.weight" in g.state_dict() and "proj_head.2.bias" in g.state_dict() and "user_emb.weight" in g.state_dict()
    ei = torch.randint(0, 120, (2, 400), device=cuda)
    aug = encoders.EdgeRemoving(0.5)(ei)
    assert aug.shape[0] == 2 and 100 < aug.shape[1] < 300
    z1, z2 = g(aug, encoders.EdgeRemoving(0.5)(ei))
    loss = losses.info_nce_loss(z1[:50], z2[:50]) + losses.info_nce_loss(z1[50:], z2[50:])
    zd1, zd2 = z1.detach().double(), z2.detach().double()

    def ref(a, b, temp=0.2):
        a, b = torch.nn.functional.normalize(a, dim=1), torch.nn.functional.normalize(b, dim=1)
        s = a @ b.T / temp
        lab = torch.arange(a.shape[0], device=a.device)
        return (torch.nn.functional.cross_entropy(s, lab) + torch.nn.functional.cross_entropy(s.T, lab)) / 2
    np.testing.assert_allclose(loss.item(), (ref(zd1[:50], zd2[:50]) + ref(zd1[50:], zd2[50:])).item(), rtol=2e-2)
    loss.backward()
    assert g.user_emb.weight.grad is not None


def test_next_batch_pairwise_semantics(cuda):
    rng = np.random.default_rng(0)
    U, I, E = 40, 60, 700
    u = rng.integers(0, U, E); i = rng.integers(0, I, E)
    data = SimpleNamespace(user_num=U, item_num=I, training_data=[[f"u{a}", f"i{b}", 1.0] for a, b in zip(u, i)],
                           user={f"u{a}": a for a in range(U)}, item={f"i{b}": b for b in range(I)})
    pos = {a: set() for a in range(U)}
    for a, b in zip(u, i):
        pos[int(a)].add(int(b))
    seen = []
    sizes = []
    for bu, bi, bj in sampling.next_batch_pairwise(data, 128):
        bu, bi, bj = bu.cpu().numpy(), bi.cpu().numpy(), bj.cpu().numpy()
        sizes.append(len(bu))
        assert len(bu) == len(bi) == len(bj)
        assert ((bj >= 0) & (bj < I)).all()
        assert all(int(j) not in pos[int(a)] for a, j in zip(bu, bj))     # negatives are never training positives
        seen += list(zip(bu.tolist(), bi.tolist()))
    assert sizes == [128] * 5 + [60]
    assert sorted(seen) == sorted(zip(u.tolist(), i.tolist()))             # one epoch = every training pair exactly once
    first = next(iter(sampling.next_batch_pairwise(data, 128)))[0].cpu().numpy()
    assert not np.array_equal(first, np.array([p[0] for p in seen[:128]]))  # reshuffled per epoch
    negs = np.concatenate([b[2].cpu().numpy() for b in data._gcf_sampler.batches(700, n_negs=1)])
    assert len(np.unique(negs)) > I // 2                                    # spread over the item set


# ------------------------------------------------------------------------------------------ social variants (cfg 4)
def _csr(z, prefix):
    shape = tuple(int(x) for x in z[f"{prefix}_shape"])
    return sp.csr_matrix((z[f"{prefix}_data"], z[f"{prefix}_indices"], z[f"{prefix}_indptr"]), shape=shape)


def test_mhcn_model_fixture(cuda, golden):
    from recommendation_b200 import social

    z = golden("mhcn_model")
    m = social.MHCNModel(int(z["user_num"]), int(z["item_num"]), z["param__user_embeddings"].shape[1], int(z["n_layers"]), float(z["ss_rate"]),
                         _csr(z, "Hs"), _csr(z, "Hj"), _csr(z, "Hp"), _csr(z, "R"))
    names = sorted(k[len("param__"):] for k in z.keys() if k.startswith("param__"))
    assert sorted(m.state_dict().keys()) == names                      # the reference's 20 keys
    m.load_state_dict({n: torch.from_numpy(z[f"param__{n}"]) for n in names})
    perms = [torch.from_numpy(p).to(cuda) for p in z["perms"]]
    out = m(z["u_idx"].tolist(), z["v_idx"].tolist(), z["neg_idx"].tolist(), perms=perms)
    for got, key in zip(out, ("batch_user", "batch_pos", "batch_neg", "ss_loss", "final_user", "final_item")):
        _close(got, z[key], rtol=1e-3, atol=1e-5)
    total = losses.bpr_loss(out[0], out[1], out[2]) + out[3]
    np.testing.assert_allclose(total.item(), float(z["rec_loss"]) + float(z["ss_loss"]), rtol=1e-4)
    total.backward()
    params = dict(m.named_parameters())
    for n in names:
        want = z[f"grad__{n}"]
        got = np.zeros_like(want) if params[n].grad is None else params[n].grad.cpu().numpy()   # sgating_*.4 are unused
        np.testing.assert_allclose(got, want, rtol=5e-3, atol=1e-5 * max(1.0, np.abs(want).max()))
    out2 = m(z["u_idx"].tolist(), z["v_idx"].tolist(), z["neg_idx"].tolist())   # training path: device randperm
    assert torch.isfinite(out2[3])


def test_diffnet_model_fixture(cuda, golden):
    from recommendation_b200 import social

    z = golden("diffnet_model")
    m = social.DiffNetModel(int(z["num_users"]), int(z["num_items"]), z["user_w"].shape[1], int(z["n_layers"]), _csr(z, "S"), _csr(z, "A"))
    assert sorted(m.state_dict().keys()) == ["item_embeddings", "user_embeddings", "weights.0", "weights.1"]
    m.load_state_dict({"user_embeddings": torch.from_numpy(z["user_w"]), "item_embeddings": torch.from_numpy(z["item_w"]),
                       "weights.0": torch.from_numpy(z["weights"][0]), "weights.1": torch.from_numpy(z["weights"][1])})
    fu = m()
    _close(fu, z["final_user"], rtol=1e-3, atol=1e-6)
    loss = m.bpr_sum_loss(fu, z["u_idx"].tolist(), z["i_idx"].tolist(), z["j_idx"].tolist(), float(z["regU"]))
    _close(loss, z["loss"], rtol=1e-4)
    loss.backward()
    _close(m.user_embeddings.grad, z["g_user_w"], atol=1e-6); _close(m.item_embeddings.grad, z["g_item_w"], atol=1e-6)
    for k in range(2):
        _close(m.weights[k].grad, z["g_weights"][k], atol=1e-6)


# ------------------------------------------------------------------------------------------ batched evaluation (8f row 2)
def test_masked_topn_vs_oracle(cuda):
    from oracle import eval_ref
    from recommendation_b200 import evaluation

    rng = np.random.default_rng(5)
    U, I, d = 300, 5000, 32
    ue = torch.from_numpy(rng.standard_normal((U, d)).astype(np.float32)).to(cuda)
    ie = torch.from_numpy(rng.standard_normal((I, d)).astype(np.float32)).to(cuda)
    tr_u = rng.integers(0, U, 9000); tr_i = rng.integers(0, I, 9000)
    tr_u[:600] = 7; tr_i[:600] = rng.permutation(I)[:600]            # a user with many rated items
    train_pos = evaluation.positives_csr(torch.from_numpy(tr_u).to(cuda), torch.from_numpy(tr_i).to(cuda), U, I)
    query = torch.from_numpy(rng.permutation(U)[:200].astype(np.int64)).to(cuda)
    for n_top, block in ((50, 1 << 30), (128, 64 * 4 * I), (1, 1 << 30)):
        idx, val = evaluation.recommend_topn(ue, ie, query, n_top, train_pos=train_pos, block_bytes=block)
        scores = (ue[query] @ ie.T).cpu().numpy()
        per_user = [np.unique(tr_i[tr_u == int(u)]) for u in query.cpu().numpy()]
        want_idx, want_val = eval_ref.masked_topn(scores, per_user, n_top)
        assert np.array_equal(idx.cpu().numpy(), want_idx)              # exact selection and order
        np.testing.assert_array_equal(val.cpu().numpy(), want_val)
        for r, its in enumerate(per_user):
            assert not set(idx[r].cpu().numpy().tolist()) & set(its.tolist())   # rated items are never recommended
    # ties: constant scores -> the lowest item ids, in order
    flat = torch.zeros(4, 1000, device=cuda)
    lib = _lib_mod.load()
    oi = torch.empty(4, 10, dtype=torch.int64, device=cuda); ov = torch.empty(4, 10, device=cuda)
    _lib_mod.check(lib.gcf_masked_topn(_lib_mod.ptr(flat), 1000, 4, 1000, None, None, None, -1e8, 10, _lib_mod.ptr(oi), _lib_mod.ptr(ov),
                                       _lib_mod.current_stream()), "topn")
    assert torch.equal(oi.cpu(), torch.arange(10).repeat(4, 1))


def test_ranking_measures_match_reference_fixture(cuda, golden):
    from recommendation_b200 import evaluation

    z = golden("eval_metrics")
    lists = torch.from_numpy(z["lists"].astype(np.int64)).to(cuda)
    U = lists.shape[0]
    ptr = z["test_ptr"]
    te_u = np.repeat(np.arange(U), np.diff(ptr))
    test_pos = evaluation.positives_csr(torch.from_numpy(te_u).to(cuda), torch.from_numpy(z["test_items"].astype(np.int64)).to(cuda), U, 80)
    got = evaluation.ranking_measures(lists, torch.arange(U, device=cuda), test_pos, [int(n) for n in z["top_ns"]])
    for row, n in zip(z["measures"], z["top_ns"]):
        m = got[int(n)]
        np.testing.assert_allclose([m["Hit Ratio"], m["Precision"], m["Recall"], m["NDCG"]], row, rtol=0, atol=1.1e-5)
    lines = evaluation.format_measures(got)
    assert lines[0] == "Top 5\n" and lines[1].startswith("Hit Ratio:") and len(lines) == 15


def test_reference_named_graph_helpers(cuda, golden):
    import scipy.sparse as sp
    from recommendation_b200 import functional as F_
    from recommendation_b200.graph import Graph, TorchGraphInterface

    g = golden("selfcf_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    ui = sp.csr_matrix((g["ui_data"], g["ui_indices"], g["ui_indptr"]), shape=(U + I, U + I))       # data.ui_adj of the reference
    norm = Graph.normalize_graph_mat(ui)                                                             # selfcf.py:240-255
    assert np.array_equal(norm.indptr, g["norm_indptr"]) and np.array_equal(norm.indices, g["norm_indices"])
    np.testing.assert_allclose(norm.data, g["norm_data"], rtol=3e-7)
    rect = ui[:U, U:]                                                                                # non-square: D^-1 A
    got = Graph.normalize_graph_mat(rect).toarray()
    rs = np.asarray(rect.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        dinv = np.where(rs > 0, 1.0 / rs, 0.0)
    np.testing.assert_allclose(got, sp.diags(dinv).dot(rect).toarray(), rtol=1e-6)
    op = TorchGraphInterface.convert_sparse_mat_to_tensor(norm)                                      # selfcf.py:219-225
    x = torch.randn(U + I, 16, device=cuda)
    np.testing.assert_allclose(F_.spmm(op, x).cpu().numpy(), norm.dot(x.cpu().numpy()), rtol=1e-4, atol=1e-5)


# ------------------------------------------------------------------------------------------ BUIR + edge dropout (8f row 3)
def test_buir_nb_fixture(cuda, golden):
    from recommendation_b200 import buir

    z = golden("buir_nb")
    U, I = int(z["n_users"]), int(z["n_items"])
    data = SimpleNamespace(user_num=U, item_num=I, norm_adj=sp.csr_matrix((z["norm_data"], z["norm_indices"], z["norm_indptr"]), shape=(U + I, U + I)))
    m = buir.BUIR_NB(data, z["online_user"].shape[1], float(z["momentum"]), int(z["n_layers"]), 0.2, drop_flag=False)
    keys = sorted(m.state_dict().keys())
    assert keys == sorted(["online_encoder.embedding_dict.user_emb", "online_encoder.embedding_dict.item_emb",
                           "target_encoder.embedding_dict.user_emb", "target_encoder.embedding_dict.item_emb", "predictor.weight", "predictor.bias"])
    assert not any(p.requires_grad for p in m.target_encoder.parameters())
    m.load_state_dict({"online_encoder.embedding_dict.user_emb": torch.from_numpy(z["online_user"]),
                       "online_encoder.embedding_dict.item_emb": torch.from_numpy(z["online_item"]),
                       "target_encoder.embedding_dict.user_emb": torch.from_numpy(z["target_user0"]),
                       "target_encoder.embedding_dict.item_emb": torch.from_numpy(z["target_item0"]),
                       "predictor.weight": torch.from_numpy(z["pred_w"]), "predictor.bias": torch.from_numpy(z["pred_b"])})
    users, items = z["users"].tolist(), z["items"].tolist()
    out = m({"user": users, "item": items})
    for got, key in zip(out, ("out_u_online", "out_u_target", "out_i_online", "out_i_target")):
        _close(got, z[key], atol=1e-5)
    loss = m.get_loss(out)
    _close(loss, z["loss"], rtol=1e-4)
    loss.backward()
    on = m.online_encoder.embedding_dict
    _close(on["user_emb"].grad, z["g_user"], atol=1e-6); _close(on["item_emb"].grad, z["g_item"], atol=1e-6)
    _close(m.predictor.weight.grad, z["g_pred_w"], atol=1e-6); _close(m.predictor.bias.grad, z["g_pred_b"], atol=1e-6)
    m.update_target(users, items)
    tg = m.target_encoder.embedding_dict
    _close(tg["user_emb"], z["target_user1"], atol=1e-6); _close(tg["item_emb"], z["target_item1"], atol=1e-6)


def test_edge_dropout_operator(cuda):
    """gcf_csr_dropout_values: mask bit-exact vs the Philox oracle, kept values rescaled, the attached transpose is the exact
    transpose of the dropped operator, and autograd through the dropped propagation matches a dense fp64 reference."""
    from oracle import philox_ref
    from recommendation_b200 import functional as F_, synth

    inter = synth.power_law_bipartite(300, 400, 6000, seed=12)
    g = CSRGraph.from_pairs(torch.from_numpy(inter.users).to(cuda), torch.from_numpy(inter.items).to(cuda), 300, 400, norm="sym", chunk=64)
    rate, seed, off = 0.3, 77, 5
    gd = g.dropout(rate, seed=seed, offset=off)
    words = philox_ref.philox4x32_10((np.arange(g.nnz, dtype=np.uint32), np.uint32(0), np.uint32(off), np.uint32(0)),
                                     (np.uint32(seed), np.uint32(0)))[0]
    keep = words < np.uint32(int((1 - rate) * 4294967296.0))
    want = np.where(keep, g.vals.cpu().numpy() * np.float32(1.0 / (1.0 - rate)), np.float32(0))
    np.testing.assert_array_equal(gd.vals.cpu().numpy(), want.astype(np.float32))
    assert abs(keep.mean() - (1 - rate)) < 0.02
    a = gd.to_scipy()
    assert abs(a - a.T).max() > 0                                   # entries are dropped independently per direction
    at = gd.transpose().to_scipy()
    assert abs(at - a.T).max() == 0                                 # the attached transpose is exact
    assert g.dropout(0.0, seed=1).vals.equal(g.vals)
    d, K = 32, 2
    x = (torch.randn(700, d, device=cuda) * 0.2).requires_grad_(True)
    w = torch.randn(700, d, device=cuda)
    (F_.propagate(gd, x, K, mode="mean") * w).sum().backward()
    ad = torch.from_numpy(a.toarray()).double()
    xd = x.detach().cpu().double().requires_grad_(True)
    layers = [xd]
    for _ in range(K):
        layers.append(ad @ layers[-1])
    (torch.stack(layers).mean(0) * w.cpu().double()).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), xd.grad.numpy(), rtol=1e-3, atol=1e-6)
    # the encoder draws a fresh mask per forward when drop_flag is set
    from recommendation_b200 import buir
    data = SimpleNamespace(user_num=300, item_num=400, norm_adj=g.to_scipy())
    enc = buir.LGCN_Encoder(data, 16, 2, 0.5, drop_flag=True)
    o1 = enc({"user": [1, 2, 3], "item": [4, 5, 6]})[0]; o2 = enc({"user": [1, 2, 3], "item": [4, 5, 6]})[0]
    assert not torch.equal(o1, o2)
    e1 = enc.get_embedding()[0]; e2 = enc.get_embedding()[0]
    assert torch.equal(e1, e2)                                      # evaluation uses the full operator


def test_lightgcn_script_surface(cuda, tmp_path):
    """load_data / evaluate / train_model of lightgcn.py:29-124 on a small text dataset."""
    import pandas as pd
    from oracle import eval_ref
    from recommendation_b200 import lightgcn as lg

    rng = np.random.default_rng(17)
    U, I = 60, 90
    pairs = np.unique(np.stack([rng.integers(0, U, 1500), rng.integers(0, I, 1500)], 1), axis=0)
    rng.shuffle(pairs)
    train, test = pairs[:1000], pairs[1000:1200]
    for name, arr in (("train.txt", train), ("test.txt", test)):
        with open(tmp_path / name, "w") as f:
            f.writelines(f"{u} {i} 1\n" for u, i in arr)
    ei, train_df, test_df, nu, ni = lg.load_data(str(tmp_path / "train.txt"), str(tmp_path / "test.txt"))
    assert ei.is_cuda and ei.shape == (2, 2 * len(train)) and nu == pairs[:1200, 0].max() + 1 and ni == pairs[:1200, 1].max() + 1
    want_ei = np.stack([np.concatenate([train[:, 0], train[:, 1] + nu]), np.concatenate([train[:, 1] + nu, train[:, 0]])])
    assert np.array_equal(ei.cpu().numpy(), want_ei)                                 # lightgcn.py:36-39, bit-exact
    ue = torch.randn(nu, 32, device=cuda); ie = torch.randn(ni, 32, device=cuda)
    train_pos = {}
    for u, i in train:
        train_pos.setdefault(int(u), set()).add(int(i))
    got = lg.evaluate(ue, ie, test_df, train_pos, k_list=[5, 10])
    want = eval_ref.lightgcn_evaluate((ue @ ie.T).cpu().numpy(), train_pos, test[:, 0], test[:, 1], [5, 10])
    for k in (5, 10):
        for m in ("HR", "P", "R", "NDCG"):
            np.testing.assert_allclose(got[k][m], want[k][m], rtol=1e-5, atol=1e-7)
    got2 = lg.evaluate(ue, ie, test_df, train_df=train_df, k_list=[10])
    assert got2[10] == got[10]
    cfg = {"embedding_dim": 32, "num_layers": 2, "optimizer": "Adam", "lr": 0.01, "weight_decay": 0.0, "n_neg": 1, "reg_weight": 1e-4,
           "loss_type": "bpr"}
    metrics = lg.train_model(cfg, str(tmp_path / "train.txt"), str(tmp_path / "test.txt"), epochs=5)
    assert set(metrics[10]) == {"HR", "P", "R", "NDCG"} and 0.0 <= metrics[10]["HR"] <= 1.0


def test_sept_encoder_and_augmentor(cuda, golden):
    from recommendation_b200 import sept

    z = golden("sept_encoder")
    U, I, K = int(z["n_users"]), int(z["n_items"]), int(z["n_layers"])
    n = U + I
    full = sp.csr_matrix((z["full_data"], z["full_indices"], z["full_indptr"]), shape=(n, n))
    dropped = sp.csr_matrix((z["drop_data"], z["drop_indices"], z["drop_indptr"]), shape=(n, n))   # produced by the reference's augmentor
    data = SimpleNamespace(user_num=U, item_num=I, norm_adj=full)
    m = sept.SEPTEncoder(data, z["emb"].shape[1], K, 0.25)
    assert sorted(m.state_dict().keys()) == ["item_embeddings.weight", "user_embeddings.weight"]
    emb = torch.from_numpy(z["emb"]).to(cuda).requires_grad_(True)
    out = m.encoder(emb, CSRGraph.from_scipy(dropped, device=cuda))                  # sept.py:220-226 on the reference's graph
    _close(out, z["out"], atol=1e-6)
    (out * torch.from_numpy(z["proj"]).to(cuda)).sum().backward()
    _close(emb.grad, z["grad"], atol=1e-6)
    # augmentor: the reference enumerates stored entries INCLUDING duplicate interactions (sept.py:53-62)
    rng = np.random.default_rng(0)
    r = rng.integers(0, 50, 3000); c = rng.integers(0, 60, 3000)
    raw = sp.coo_matrix((np.ones(3000, np.float32), (r, c)), shape=(50, 60))          # many duplicates
    a1 = sept.GraphAugmentor.edge_dropout(raw, 0.3)
    a2 = sept.GraphAugmentor.edge_dropout(raw, 0.3)
    for a in (a1, a2):
        kept = a.to_scipy()
        assert abs(kept.sum() - int(3000 * 0.7)) < 1e-3                                  # exactly int(n (1 - rate)) entries kept
        assert (raw.tocsr() - kept).min() >= 0                                          # a sub-multiset of the stored entries
        assert np.allclose(kept.data, np.round(kept.data))                             # 0/1 entries, duplicates summed
    assert abs(a1.to_scipy() - a2.to_scipy()).sum() > 0                               # a fresh draw per call
    m.train(); u1, v1 = m()
    m.eval(); u2, v2 = m(); u3, v3 = m()
    assert u1.shape == (U, z["emb"].shape[1]) and torch.equal(u2, u3) and not torch.equal(u1, u2)


def test_gpu_kmeans_against_sklearn_lloyd(cuda):
    """Same initial centroids -> the same Lloyd fixed point as scikit-learn (full-batch, 25 iterations), plus the invariants
    of the result and NCL's k cap (ncl.py:347-356)."""
    from sklearn.cluster import KMeans
    from recommendation_b200 import kmeans as km

    rng = np.random.default_rng(3)
    centers = rng.standard_normal((12, 16)) * 4
    x = (centers[rng.integers(0, 12, 4000)] + rng.standard_normal((4000, 16)) * 0.7).astype(np.float32)
    init = x[rng.permutation(4000)[:12]].copy()
    c, idx, obj = km.kmeans(torch.from_numpy(x).to(cuda), 12, niter=25, init=torch.from_numpy(init))
    ref = KMeans(n_clusters=12, init=init, n_init=1, max_iter=25, tol=0.0, algorithm="lloyd").fit(x)
    np.testing.assert_allclose(obj, ref.inertia_, rtol=1e-3)
    assert (idx.cpu().numpy() == ref.labels_).mean() > 0.995
    # invariants: every point sits with its nearest centroid; every centroid is the mean of its points
    d2 = ((x[:, None, :] - c.cpu().numpy()[None]) ** 2).sum(-1)
    assert (d2.argmin(1) == idx.cpu().numpy()).mean() > 0.999
    c2, idx2, _ = km.kmeans(torch.from_numpy(x).to(cuda), 12, niter=60, init=torch.from_numpy(init))   # converged
    for j in range(12):
        np.testing.assert_allclose(c2[j].cpu().numpy(), x[idx2.cpu().numpy() == j].mean(0), rtol=1e-3, atol=1e-3)
    cc, ii, k_used = km.run_kmeans(torch.from_numpy(x[:400]).to(cuda), 1000)
    assert k_used == max(2, 400 // 39) and cc.shape == (k_used, 16) and ii.shape == (400,) and ii.dtype == torch.int64
    # an empty cluster (duplicate initial centroids) is re-seeded instead of producing NaNs
    bad = np.repeat(init[:1], 12, 0)
    c3, idx3, _ = km.kmeans(torch.from_numpy(x).to(cuda), 12, niter=25, init=torch.from_numpy(bad))
    assert torch.isfinite(c3).all() and len(torch.unique(idx3)) > 1


def test_ncl_training_iteration_runs_and_learns(cuda):
    """ncl.py:308-329 end to end on a small graph: E-step (GPU k-means), batch sampler, BPR + ssl_layer + ProtoNCE, Adam."""
    from recommendation_b200 import ncl as ncl_mod
    from recommendation_b200.losses import NCLLosses

    rng = np.random.default_rng(2)
    U, I, E = 400, 500, 12000
    u = rng.integers(0, U, E); i = (rng.zipf(1.5, E) - 1) % I
    raw = sp.coo_matrix((np.ones(2 * E, np.float32), (np.concatenate([u, i + U]), np.concatenate([i + U, u]))), shape=(U + I, U + I))
    # The learning check runs on the sym-normalised operator: with ncl.py's RAW adjacency (ncl.py:76-85, covered by the
    # fixtures) activations grow by ~deg per layer (score differences ~1e5 here), the eps-sigmoid saturates and the
    # trajectory is chaotic -- an fp32 CPU restatement of this loop moves 0.67 -> 0.26 in 4 epochs on the normalised one.
    deg = np.asarray(raw.tocsr().sum(1)).ravel()
    dinv = np.where(deg > 0, 1.0 / np.sqrt(np.maximum(deg, 1e-12)), 0.0)
    norm = (sp.diags(dinv) @ raw.tocsr() @ sp.diags(dinv)).tocoo().astype(np.float32)
    data = SimpleNamespace(user_num=U, item_num=I, norm_adj=norm, training_data=[[f"u{a}", f"i{b}", 1.0] for a, b in zip(u, i)],
                           user={f"u{a}": a for a in range(U)}, item={f"i{b}": b for b in range(I)})
    torch.manual_seed(0)
    model = encoders.LGCNEncoder(data, 64, 3)
    ncl = NCLLosses(U, I, 0.1, 1e-6, 1.5, 8e-8, 512)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    k = ncl_mod.e_step(model, ncl, 20)
    assert k == max(2, min(20, U // 39)) and ncl.user_2cluster.shape == (U,) and ncl.item_centroids.shape[1] == 64
    per_epoch = []
    for epoch in range(4):
        rec = []
        for n, batch in enumerate(sampling.next_batch_pairwise(data, 512)):
            total, parts = ncl_mod.ncl_step(model, ncl, opt, batch, 1e-4, 512, 1, k=k, refresh_clusters=(n % 8 == 0))
            assert torch.isfinite(total) and all(torch.isfinite(v) for v in parts.values())
            rec.append(parts["rec"].item())
        per_epoch.append(float(np.mean(rec)))
    assert 0.6 < per_epoch[0] < 0.72, per_epoch         # -log(sigmoid(~0)) at the start
    assert per_epoch[-1] < 0.75 * per_epoch[0], per_epoch      # the ranking loss goes down (epoch means: single batches are noisy)
    # the reference's raw operator: one epoch runs and stays finite (values are not meaningful in that regime, see above)
    data_raw = SimpleNamespace(user_num=U, item_num=I, norm_adj=raw, training_data=data.training_data, user=data.user, item=data.item)
    model_raw = encoders.LGCNEncoder(data_raw, 64, 3)
    opt_raw = torch.optim.Adam(model_raw.parameters(), lr=1e-3)
    ncl_mod.e_step(model_raw, ncl, 20)
    for n, batch in enumerate(sampling.next_batch_pairwise(data_raw, 512)):
        total, parts = ncl_mod.ncl_step(model_raw, ncl, opt_raw, batch, 1e-4, 512, 1, k=k, refresh_clusters=(n % 8 == 0))
        assert torch.isfinite(total) and all(torch.isfinite(v) for v in parts.values())


# ------------------------------------------------------------------------------------------ MHCN motif matrices (8f row 4)
def _rand_sparse(rng, n_rows, n_cols, nnz, integer=True):
    r, c = rng.integers(0, n_rows, nnz), rng.integers(0, n_cols, nnz)
    v = rng.integers(1, 4, nnz).astype(np.float32) if integer else rng.standard_normal(nnz).astype(np.float32)
    m = sp.coo_matrix((v, (r, c)), shape=(n_rows, n_cols)).tocsr()
    m.sum_duplicates(); m.sort_indices()
    return m


def _to_scipy(g):
    m = sp.csr_matrix((g.vals.cpu().numpy(), g.col_idx.cpu().numpy(), g.row_ptr.cpu().numpy()), shape=(g.n_rows, g.n_cols))
    m.sort_indices()
    return m


@pytest.mark.parametrize("shape", [(60, 45, 70, 400), (500, 300, 400, 9000), (7, 5, 9, 0)])
def test_sparse_product_kernels_vs_scipy(cuda, shape):
    """gcf_csr_sample, gcf_spgemm_masked and the expand-sort-compress product against scipy (integer-valued operands:
    bit-exact values, identical patterns; rectangular shapes, empty rows, an empty operand)."""
    from recommendation_b200 import motifs

    m, k, n, nnz = shape
    rng = np.random.default_rng(m)
    A, B, M = _rand_sparse(rng, m, k, nnz), _rand_sparse(rng, k, n, nnz), _rand_sparse(rng, m, n, nnz)
    gA, gB, gM = (CSRGraph.from_scipy(x, device=cuda) for x in (A, B, M))
    # masked product on M's pattern
    got = motifs.masked_product(gA, gB.transpose(), gM).cpu().numpy()
    want = np.asarray((A @ B).multiply(M).todense())
    dense = np.zeros((m, n), np.float32)
    rows = motifs.entry_rows(gM).cpu().numpy()
    dense[rows, gM.col_idx.cpu().numpy()] = got
    assert np.array_equal(dense, want)
    # sampling one matrix at another's pattern
    X = _rand_sparse(rng, m, n, nnz)
    s = motifs.csr_sample(CSRGraph.from_scipy(X, device=cuda), gM).cpu().numpy()
    assert np.array_equal(s, np.asarray(X.todense())[rows, gM.col_idx.cpu().numpy()])
    # full product, whole and in forced row blocks (budget = the largest single row, so every split is at a row boundary)
    W = (A @ B).tocsr(); W.sort_indices()
    per_row = np.asarray((A != 0).astype(np.int64) @ np.diff(B.indptr).astype(np.int64)).ravel() if nnz else np.zeros(1, np.int64)
    for max_products in (1 << 27, int(per_row.max()) + 5):
        C = _to_scipy(motifs.spgemm(gA, gB, max_products=max_products))
        assert np.array_equal(C.indptr, W.indptr) and np.array_equal(C.indices, W.indices) and np.array_equal(C.data, W.data)
    if per_row.max() > 1:   # a single row above the budget is reported, not truncated
        with pytest.raises(RuntimeError, match="max_products"):
            motifs.spgemm(gA, gB, max_products=1)


def test_hyper_adj_mats_match_reference_fixture(cuda, golden):
    """motifs.build_hyper_adj_mats on the GPU against the reference's own build_hyper_adj_mats (mhcn.py:340-368) on the same
    S / Y: identical sparsity patterns, values within 2 ulp; and the matrices drive MHCNModel."""
    from recommendation_b200 import motifs, social

    z = golden("mhcn_motifs")
    S, Y = _csr(z, "S"), _csr(z, "Y")
    got = motifs.build_hyper_adj_mats(CSRGraph.from_scipy(S, device=cuda), CSRGraph.from_scipy(Y, device=cuda))
    for g, name in zip(got, ("Hs", "Hj", "Hp")):
        want = _csr(z, name); want.eliminate_zeros(); want.sort_indices()
        h = _to_scipy(g)
        assert h.nnz == want.nnz > 0, name
        assert np.array_equal(h.indptr, want.indptr) and np.array_equal(h.indices, want.indices), name   # bit-exact structure
        np.testing.assert_allclose(h.data, want.data, rtol=3e-7, atol=0)
    R = CSRGraph.from_scipy(Y, norm="row", device=cuda)
    model = social.MHCNModel(S.shape[0], Y.shape[1], 16, 2, 0.01, got[0], got[1], got[2], R)
    out = model([0, 1, 2, 3], [0, 1, 2, 3], [4, 5, 6, 7])
    assert all(torch.isfinite(o).all() for o in out)


# ------------------------------------------------------------------------------------------ sept_social.py (8f row 3)
def test_sept_social_fixture(cuda, golden):
    """SEPTSocial against the reference's own SEPT (sept_social.py) on the same parameters / batch: view matrices (identical
    pattern), encoder outputs, predictions, pseudo-labels, losses and parameter gradients."""
    from recommendation_b200 import sept_social

    z = golden("sept_social")
    U, I = int(z["user_num"]), int(z["item_num"])
    adj = sp.coo_matrix((z["adj_data"], (z["adj_row"], z["adj_col"])), shape=(U + I, U + I))   # raw, duplicates kept
    data = SimpleNamespace(user_num=U, item_num=I, norm_adj=adj, interaction_mat=_csr(z, "Y"))
    m = sept_social.SEPTSocial(data, _csr(z, "bi"), emb_size=z["user_w"].shape[1], n_layers=int(z["n_layers"]),
                               ss_rate=float(z["ss_rate"]), ins_cnt=int(z["ins_cnt"]), reg=float(z["reg"]))
    assert sorted(m.state_dict().keys()) == ["item_embeddings", "user_embeddings"]
    m.load_state_dict({"user_embeddings": torch.from_numpy(z["user_w"]), "item_embeddings": torch.from_numpy(z["item_w"])})
    m.build()
    for g, name in ((m.social_mat, "social"), (m.sharing_mat, "sharing")):
        want = _csr(z, name); want.sort_indices()
        h = _to_scipy(g)
        assert np.array_equal(h.indptr, want.indptr) and np.array_equal(h.indices, want.indices), name
        np.testing.assert_allclose(h.data, want.data, rtol=5e-7)
    labels = tuple(z[k] for k in ("f_pos", "sh_pos", "r_pos"))
    rec_loss, nd, total = m.iteration_losses(z["user_idx"], z["pos_idx"], z["neg_idx"], labels=labels)
    _close(m.rec_user_embeddings, z["rec_user"]); _close(m.rec_item_embeddings, z["rec_item"])
    _close(m.sharing_view_embeddings, z["sharing_view"]); _close(m.friend_view_embeddings, z["friend_view"])
    for got, key in zip(m.last_predictions, ("social_prediction", "sharing_prediction", "rec_prediction")):
        _close(got, z[key], rtol=1e-3, atol=1e-6)
    for got, key, mk in zip(m.last_labels, ("f_pos", "sh_pos", "r_pos"), ("f_margin", "sh_margin", "r_margin")):
        clear = z[mk] > 1e-5                               # rows whose K-th and (K+1)-th probabilities are not fp32-close
        assert clear.mean() > 0.9
        assert np.array_equal(np.sort(got.cpu().numpy()[clear], 1), np.sort(z[key][clear], 1)), key
    np.testing.assert_allclose(rec_loss.item(), float(z["rec_loss"]), rtol=1e-3)
    # the B_u x B_u denominators run on bf16 tensor-core logits: north-star tolerance 2e-2
    np.testing.assert_allclose(nd.item(), float(z["nd_f"]) + float(z["nd_s"]) + float(z["nd_r"]), rtol=2e-2)
    np.testing.assert_allclose(total.item(), float(z["total"]), rtol=2e-2)
    total.backward()
    for p, key in ((m.user_embeddings, "g_user"), (m.item_embeddings, "g_item")):
        got, want = p.grad.cpu().numpy(), z[key]
        assert np.abs(got - want).max() <= 2e-2 * np.abs(want).max() + 1e-7, key
        assert np.abs(got - want).sum() <= 1e-2 * np.abs(want).sum(), key
    # training path: own pseudo-labels, augmented operator from the edge-dropout augmentor
    from recommendation_b200.sept import GraphAugmentor
    aug = GraphAugmentor.edge_dropout(adj, 0.3, seed=3)
    out = m.iteration_losses(z["user_idx"], z["pos_idx"], z["neg_idx"], aug_adj=aug)
    assert all(torch.isfinite(t) for t in out)


# ------------------------------------------------------------------------------------------ esrf.py (8f row 3)
def test_esrf_fixture(cuda, golden):
    """recommendation_b200.esrf against the reference's own ESRF pieces: motif adjacency (identical pattern), joint adjacency,
    Generator (replayed uniform draws), Discriminator in both modes, the loss lines of trainModel and their gradients."""
    from recommendation_b200 import esrf

    z = golden("esrf")
    K, seg, regU, beta = int(z["K"]), int(z["segment"]), float(z["regU"]), float(z["beta"])
    U, I = _csr(z, "Y").shape
    gS, gY = CSRGraph.from_scipy(_csr(z, "S"), device=cuda), CSRGraph.from_scipy(_csr(z, "Y"), device=cuda)
    A = esrf.build_motif_induced_adjacency_matrix(gS, gY)
    want = _csr(z, "A"); want.eliminate_zeros(); want.sort_indices()
    h = _to_scipy(A)
    assert np.array_equal(h.indptr, want.indptr) and np.array_equal(h.indices, want.indices)      # bit-exact structure
    np.testing.assert_allclose(h.data, want.data, rtol=3e-7)
    joint = esrf.create_joint_sparse_adjacency(torch.from_numpy(z["users"]).to(cuda), torch.from_numpy(z["items"]).to(cuda), U, I)
    wj = _csr(z, "joint"); wj.sort_indices()
    hj = _to_scipy(joint)
    assert np.array_equal(hj.indptr, wj.indptr) and np.array_equal(hj.indices, wj.indices)
    np.testing.assert_allclose(hj.data, wj.data, rtol=3e-7)

    d = z["gen_relation"].shape[1]
    gen = esrf.Generator(U, d, int(z["n_layers_G"]), K)
    dis = esrf.Discriminator(U, I, d, int(z["n_layers_D"]))
    assert sorted(gen.state_dict().keys()) == list(z["gen_state_keys"]) and sorted(dis.state_dict().keys()) == list(z["dis_state_keys"])
    with torch.no_grad():
        gen.relation_embeddings.copy_(torch.from_numpy(z["gen_relation"])); gen.projection_head.copy_(torch.from_numpy(z["gen_projection"]))
        gen.c_selector.copy_(torch.from_numpy(z["gen_selector"]))
        dis.user_embeddings.copy_(torch.from_numpy(z["dis_user"])); dis.item_embeddings.copy_(torch.from_numpy(z["dis_item"]))
    alt = gen(A, seg, noise=torch.from_numpy(z["noise"]).to(cuda))
    _close(alt, z["alt"], rtol=2e-3, atol=1e-6)
    # pretraining pass
    pu, pi = dis(joint, None, False, 0, K)
    _close(pu, z["pre_user"]); _close(pi, z["pre_item"])
    pair, reg = esrf.pairwise_losses(pu, pi, z["user_idx"], z["i_idx"], z["j_idx"], regU)
    np.testing.assert_allclose([pair.item(), reg.item()], [float(z["pre_pair"]), float(z["pre_reg"])], rtol=1e-3)
    g = torch.autograd.grad(pair + reg, (dis.user_embeddings, dis.item_embeddings))
    for got, key in zip(g, ("g_pre_user", "g_pre_item")):
        np.testing.assert_allclose(got.cpu().numpy(), z[key], rtol=5e-3, atol=1e-5 * np.abs(z[key]).max())
    # adversarial pass
    su, si = dis(joint, alt, True, 0, K)
    _close(su, z["soc_user"]); _close(si, z["soc_item"])
    pair, reg = esrf.pairwise_losses(su, si, z["user_idx"], z["i_idx"], z["j_idx"], regU)
    adv, g_adv = esrf.adversarial_losses(su, si, alt, z["user_idx"], z["i_idx"], K)
    np.testing.assert_allclose([pair.item(), reg.item(), adv.item(), beta * g_adv.item()],
                               [float(z["pair"]), float(z["reg"]), float(z["adv"]), float(z["g_loss"])], rtol=1e-3)
    gd = torch.autograd.grad(pair + reg + beta * adv, (dis.user_embeddings, dis.item_embeddings), retain_graph=True)
    for got, key in zip(gd, ("g_d_user", "g_d_item")):
        np.testing.assert_allclose(got.cpu().numpy(), z[key], rtol=5e-3, atol=1e-5 * np.abs(z[key]).max())
    gg = torch.autograd.grad(beta * g_adv, (gen.relation_embeddings, gen.c_selector))
    for got, key in zip(gg, ("g_g_relation", "g_g_selector")):
        np.testing.assert_allclose(got.cpu().numpy(), z[key], rtol=1e-2, atol=1e-4 * np.abs(z[key]).max())
    assert torch.isfinite(gen(A, 0)).all()      # training path: own noise


# ------------------------------------------------------------------------------------------ r02 reference-generated fixtures
def test_dnn_encoder_reference_fixture(cuda, golden):
    """ssl4rec.py:162-196,221-224: the reference's own DNNEncoder (eval mode), its forward, cal_cl_loss, the training-loop
    loss and every parameter gradient (tests/golden/make_golden_r02.py)."""
    z = golden("dnn_encoder")
    data = SimpleNamespace(user_num=int(z["n_users"]), item_num=int(z["n_items"]))
    m = encoders.DNNEncoder(data, int(z["emb_size"]), float(z["drop_rate"]), float(z["tau"]), int(z["n_layers"]))
    names = [k[len("param."):] for k in z if k.startswith("param.")]
    assert sorted(names) == sorted(m.state_dict().keys())                 # state_dict keys are part of the interface
    m.load_state_dict({n: torch.from_numpy(z[f"param.{n}"]) for n in names})
    m.eval()
    u, i = z["u"].tolist(), z["i"].tolist()
    q, k = m(u, i)
    np.testing.assert_allclose(q.detach().cpu().numpy(), z["q"], rtol=1e-3, atol=1e-5)   # fp32 path (gather + cuBLAS towers)
    np.testing.assert_allclose(k.detach().cpu().numpy(), z["k"], rtol=1e-3, atol=1e-5)
    cl = m.cal_cl_loss(i)
    rec = losses.batch_softmax_loss(q, k, float(z["tau"]))
    total = rec + losses.l2_reg_loss(1e-4, q, k) + 0.1 * cl
    np.testing.assert_allclose(cl.item(), float(z["cl"]), rtol=2e-2)                       # bf16 logits
    np.testing.assert_allclose(rec.item(), float(z["rec"]), rtol=2e-2)
    np.testing.assert_allclose(total.item(), float(z["total"]), rtol=2e-2)
    total.backward()
    for n, p in m.named_parameters():
        want = z[f"grad.{n}"]
        assert p.grad is not None, n
        # bf16 logits: 2e-2 on each element or, for entries that are sums over the batch with cancellation (biases), 4e-2 of
        # the gradient's own scale
        np.testing.assert_allclose(p.grad.cpu().numpy(), want, rtol=2e-2, atol=4e-2 * np.abs(want).max() + 1e-7, err_msg=n)


def test_evaluate_model_counts_cold_test_items(cuda, golden):
    """ncl.py:253-277 + 133-178 on the reference's own Interaction: test items that never occur in training stay in
    test_set, so they count in len(origin[u]) (Hit Ratio / Recall denominators, ideal DCG) although they cannot be hit."""
    from recommendation_b200 import evaluation

    z = golden("eval_cold")
    tr_u, tr_i = z["train_users"].tolist(), z["train_items"].tolist()
    users, items = sorted(set(tr_u)), sorted(set(tr_i))                                   # Interaction._build, ncl.py:55-61
    data = SimpleNamespace(user={u: k for k, u in enumerate(users)}, item={i: k for k, i in enumerate(items)},
                           user_num=len(users), item_num=len(items), training_data=list(zip(tr_u, tr_i)), test_set={})
    for u, i in zip(z["test_users"].tolist(), z["test_items"].tolist()):
        data.test_set.setdefault(u, {})[i] = 1
    assert any(i not in data.item for its in data.test_set.values() for i in its)         # the fixture does hold cold items
    got = evaluation.evaluate_model(torch.from_numpy(z["user_emb"]).to(cuda), torch.from_numpy(z["item_emb"]).to(cuda), data,
                                    [int(n) for n in z["top_ns"]])
    want = [str(s) for s in z["strings"]]
    assert len(got) == len(want)
    for g, w in zip(got, want):
        if ":" in w:
            assert g.split(":")[0] == w.split(":")[0]
            np.testing.assert_allclose(float(g.split(":")[1]), float(w.split(":")[1]), rtol=0, atol=1.1e-5)
        else:
            assert g == w
