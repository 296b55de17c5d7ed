"""Pin the CPU oracle against outputs of the reference itself (tests/golden/*.npz, see make_golden.py) and against
Random123's known-answer vectors.  CPU only."""
import numpy as np
import pytest
import torch

from oracle import graph_ref, losses_ref, lightgcn_ref, philox_ref

ULP2 = 2.4e-7  # 2 ulp of fp32: numpy power(-0.5) vs 1/sqrt disagree by <= 1 ulp each


def T(a, grad=False):
    t = torch.from_numpy(np.asarray(a)).clone()
    if grad:
        t.requires_grad_(True)
    return t


# ---------------------------------------------------------------- graph build
def test_sym_graph_matches_selfcf(golden):
    g = golden("selfcf_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    ei = graph_ref.bipartite_edge_index(g["u_idx"], g["i_idx"], U)
    rp, ci, mult = graph_ref.coo_to_canonical_csr(ei[0], ei[1], None, U + I, U + I)
    assert np.array_equal(rp, g["ui_indptr"])
    assert np.array_equal(ci, g["ui_indices"])
    assert np.array_equal(mult, g["ui_data"])  # duplicates summed: exact small integers
    assert mult.max() >= 2, "fixture must contain duplicate interactions"
    vals, rowsum, dinv = graph_ref.normalize_csr(rp, ci, mult, U + I, U + I, "sym")
    assert np.array_equal(rowsum, g["rowsum"])
    assert np.array_equal(rp, g["norm_indptr"]) and np.array_equal(ci, g["norm_indices"])
    np.testing.assert_allclose(vals, g["norm_data"], rtol=ULP2, atol=0)


def test_sym_graph_matches_ssl4rec(golden):
    g = golden("ssl4rec_losses")
    U, I = int(g["n_users"]), int(g["n_items"])
    ei = graph_ref.bipartite_edge_index(g["u_idx"], g["i_idx"], U)
    rp, ci, mult = graph_ref.coo_to_canonical_csr(ei[0], ei[1], None, U + I, U + I)
    vals, _, _ = graph_ref.normalize_csr(rp, ci, mult, U + I, U + I, "sym")
    assert np.array_equal(rp, g["norm_indptr"]) and np.array_equal(ci, g["norm_indices"])
    np.testing.assert_allclose(vals, g["norm_data"], rtol=ULP2, atol=0)


def test_raw_graph_matches_ncl(golden):
    g = golden("ncl_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    # reference: interleaved (u,i),(i,u) insertion order with duplicates kept (ncl.py:76-85)
    ref = graph_ref.coo_to_canonical_csr(g["coo_row"], g["coo_col"], g["coo_data"], U + I, U + I)
    ei = graph_ref.bipartite_edge_index(g["u_idx"], g["i_idx"], U)
    ours = graph_ref.coo_to_canonical_csr(ei[0], ei[1], None, U + I, U + I)
    for a, b in zip(ref, ours):
        assert np.array_equal(a, b)
    assert np.array_equal(graph_ref.degrees(ei[0], U + I), graph_ref.degrees(g["coo_row"], U + I))


def test_gcn_norm_equals_sym_normalisation(golden):
    g = golden("selfcf_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    ei = graph_ref.bipartite_edge_index(g["u_idx"], g["i_idx"], U)
    w, deg = graph_ref.gcn_norm_weights(ei, U + I)
    assert np.array_equal(deg.astype(np.float32), g["rowsum"])
    rp, ci, vals = graph_ref.coo_to_canonical_csr(ei[1], ei[0], w, U + I, U + I)  # M[col,row] = w, duplicates summed
    np.testing.assert_allclose(vals, g["norm_data"], rtol=4 * ULP2, atol=0)


def test_transpose_roundtrip():
    rng = np.random.default_rng(0)
    r, c = rng.integers(0, 30, 200), rng.integers(0, 45, 200)
    rp, ci, v = graph_ref.coo_to_canonical_csr(r, c, rng.random(200).astype(np.float32), 30, 45)
    t = graph_ref.csr_transpose(rp, ci, v, 30, 45)
    tt = graph_ref.csr_transpose(*t, 45, 30)
    for a, b in zip((rp, ci, v), tt):
        assert np.array_equal(a, b)


# ---------------------------------------------------------------- propagation
def _sym_csr(g):
    U, I = int(g["n_users"]), int(g["n_items"])
    return g["norm_indptr"], g["norm_indices"], g["norm_data"], U, I


def test_propagate_mean_matches_selfcf_encoder(golden):
    g, e = golden("selfcf_graph"), golden("selfcf_encoder")
    rp, ci, v, U, I = _sym_csr(g)
    x0 = np.concatenate([e["user_w"], e["item_w"]])
    K = int(e["n_layers"])
    _, final = graph_ref.propagate(rp, ci, v, x0, K, "mean")
    np.testing.assert_allclose(final[:U], e["user_all"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(final[U:], e["item_all"], rtol=1e-5, atol=1e-6)
    # backward: A symmetric => dL/dE0 = mean_k A^k proj
    proj = np.concatenate([e["proj_u"], e["proj_i"]])
    _, gfinal = graph_ref.propagate(rp, ci, v, proj, K, "mean")
    np.testing.assert_allclose(gfinal[:U], e["grad_user_w"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(gfinal[U:], e["grad_item_w"], rtol=1e-4, atol=1e-5)


def test_propagate_raw_matches_ncl_encoder(golden):
    g, e = golden("ncl_graph"), golden("ncl_encoder")
    U, I = int(g["n_users"]), int(g["n_items"])
    rp, ci, v = graph_ref.coo_to_canonical_csr(g["coo_row"], g["coo_col"], g["coo_data"], U + I, U + I)
    x0 = np.concatenate([e["user_w"], e["item_w"]])
    K = int(e["n_layers"])
    layers, final = graph_ref.propagate(rp, ci, v, x0, K, "mean")
    np.testing.assert_allclose(final[:U], e["user_out"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(final[U:], e["item_out"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(np.stack(layers), e["all_emb"], rtol=1e-5, atol=1e-5)


def test_lgconv_matches_selfcf_fixture(golden):
    """The restated PyG LGConv (lightgcn.py cannot be imported) equals the importable reference encoder on the same
    graph: sum over layers == (K+1) * mean over layers."""
    g, e = golden("selfcf_graph"), golden("selfcf_encoder")
    U, I = int(g["n_users"]), int(g["n_items"])
    K = int(e["n_layers"])
    ei = torch.from_numpy(graph_ref.bipartite_edge_index(g["u_idx"], g["i_idx"], U))
    uw, iw = T(e["user_w"], True), T(e["item_w"], True)
    ue, ie = lightgcn_ref.lightgcn_forward(uw, iw, ei, K)
    np.testing.assert_allclose(ue.detach().numpy() / (K + 1), e["user_all"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(ie.detach().numpy() / (K + 1), e["item_all"], rtol=1e-5, atol=1e-6)
    loss = ((ue * T(e["proj_u"])).sum() + (ie * T(e["proj_i"])).sum()) / (K + 1)
    loss.backward()
    np.testing.assert_allclose(uw.grad.numpy(), e["grad_user_w"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(iw.grad.numpy(), e["grad_item_w"], rtol=1e-4, atol=1e-5)


# ---------------------------------------------------------------- losses
def _check(loss, params, want_loss, want_grads, rtol=1e-5, atol=1e-6):
    np.testing.assert_allclose(loss.detach().numpy(), want_loss, rtol=rtol, atol=atol)
    grads = torch.autograd.grad(loss, params, allow_unused=True)
    for g, w in zip(grads, want_grads):
        g = np.zeros_like(w) if g is None else g.numpy()
        np.testing.assert_allclose(g, w, rtol=rtol * 10, atol=atol)


def test_ncl_losses(golden):
    z = golden("ncl_losses")
    ue, pe, ne = T(z["ue"], True), T(z["pe"], True), T(z["ne"], True)
    _check(losses_ref.bpr_log_eps_sigmoid(ue, pe, ne), (ue, pe, ne), z["bpr"], (z["g_bpr_u"], z["g_bpr_p"], z["g_bpr_n"]))
    _check(losses_ref.l2_reg(1e-3, ue, pe, ne), (ue, pe, ne), z["reg"], (z["g_reg_u"], z["g_reg_p"], z["g_reg_n"]))
    v1, v2 = T(z["v1"], True), T(z["v2"], True)
    _check(losses_ref.info_nce(v1, v2, 0.2), (v1, v2), z["nce"], (z["g_nce_1"], z["g_nce_2"]))
    _check(losses_ref.info_nce(v1 * 0.3, v2 * 0.3, 0.5, cos=False), (v1, v2), z["nce_nocos"], (z["g_nce_nocos_1"], z["g_nce_nocos_2"]))
    ctx, ini = T(z["ctx"], True), T(z["ini"], True)
    nU = int(z["n_users"])
    l = losses_ref.ssl_layer_loss(ctx, ini, T(z["bu"]), T(z["bp"]), nU, float(z["ssl_temp"]), float(z["ssl_reg"]), float(z["alpha"]))
    _check(l, (ctx, ini), z["ssl"], (z["g_ssl_ctx"], z["g_ssl_ini"]), rtol=1e-4, atol=1e-9)
    l = losses_ref.proto_nce(ini, T(z["bu"]), T(z["bp"]), nU, T(z["user_centroids"]), T(z["user_2cluster"]), T(z["item_centroids"]),
                             T(z["item_2cluster"]), float(z["ssl_temp"]), float(z["proto_reg"]), int(z["batch_size"]))
    _check(l, (ini,), z["proto"], (z["g_proto_ini"],), rtol=1e-4, atol=1e-9)


def test_directau_losses(golden):
    z = golden("directau_losses")
    xu, xp, xn = T(z["xu"], True), T(z["xp"], True), T(z["xn"], True)
    gamma = float(z["gamma"])
    _check(losses_ref.alignment(xu, xp), (xu, xp), z["align"], (z["g_align_u"], z["g_align_p"]))
    _check(losses_ref.uniformity(xu), (xu,), z["unif"], (z["g_unif"],), rtol=1e-4)
    _check(losses_ref.directau_loss(xu, xp, gamma), (xu, xp), z["calc"], (z["g_calc_u"], z["g_calc_p"]), rtol=1e-4)
    train = (losses_ref.directau_loss(xu, xp, gamma) - losses_ref.directau_loss(xu, xn, gamma)
             + losses_ref.l2_reg(float(z["reg"]), xu, xp, xn) / int(z["batch_size"]))
    _check(train, (xu, xp, xn), z["train"], (z["g_train_u"], z["g_train_p"], z["g_train_n"]), rtol=1e-4)


def test_ssl4rec_and_gcl_losses(golden):
    z = golden("ssl4rec_losses")
    a, b = T(z["a"], True), T(z["b"], True)
    _check(losses_ref.batch_softmax(a, b, 0.2), (a, b), z["batch_softmax"], (z["g_bs_a"], z["g_bs_b"]))
    _check(losses_ref.info_nce(a, b, 0.15), (a, b), z["nce"], (z["g_nce_a"], z["g_nce_b"]))
    z = golden("gcl_losses")
    z1, z2 = T(z["z1"], True), T(z["z2"], True)
    _check(losses_ref.info_nce_symmetric(z1, z2, 0.2), (z1, z2), z["info_nce"], (z["g_z1"], z["g_z2"]))
    ue, pe, ne = T(z["ue"], True), T(z["pe"], True), T(z["ne"], True)
    _check(losses_ref.bpr_gcl(ue, pe, ne, float(z["reg_weight"])), (ue, pe, ne), z["bpr_reg"], (z["g_u"], z["g_p"], z["g_n"]))


def test_selfcf_he(golden):
    z = golden("selfcf_he")
    g = golden("selfcf_graph")
    rp, ci, v, U, I = _sym_csr(g)
    uw, iw = T(z["user_w"], True), T(z["item_w"], True)
    W, bias = T(z["pred_w"], True), T(z["pred_b"], True)
    m = float(z["momentum"])
    K = int(z["n_layers"])
    # encoder through autograd: dense A (tiny graph)
    A = torch.zeros(U + I, U + I)
    row_of = np.repeat(np.arange(U + I), np.diff(rp))
    A[torch.from_numpy(row_of), torch.from_numpy(ci.astype(np.int64))] = torch.from_numpy(v)
    x = torch.cat([uw, iw]); layers = [x]
    for _ in range(K):
        layers.append(A @ layers[-1])
    final = torch.stack(layers).mean(0)
    uo, io = final[:U], final[U:]
    users, items = T(z["users"]), T(z["items"])
    t_u = T(z["his_u0"])[users] * m + uo[users].detach() * (1 - m)
    t_i = T(z["his_i0"])[items] * m + io[items].detach() * (1 - m)
    p_u, p_i = uo[users] @ W.T + bias, io[items] @ W.T + bias
    for got, want in ((p_u, z["p_u"]), (t_u, z["t_u"]), (p_i, z["p_i"]), (t_i, z["t_i"])):
        np.testing.assert_allclose(got.detach().numpy(), want, rtol=1e-5, atol=1e-6)
    loss = losses_ref.selfcf_loss(p_u, t_u, p_i, t_i)
    _check(loss, (uw, iw, W, bias), z["loss"], (z["g_user_w"], z["g_item_w"], z["g_pred_w"], z["g_pred_b"]), rtol=1e-4)
    his_u1 = T(z["his_u0"]).clone(); his_u1[users] = uo[users].detach()
    np.testing.assert_allclose(his_u1.numpy(), z["his_u1"], rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------- Philox
def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    def run(c, k):
        return [int(x) for x in philox_ref.philox4x32_10([np.uint32(v) for v in c], [np.uint32(v) for v in k])]
    assert run([0, 0, 0, 0], [0, 0]) == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]
    assert run([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2) == [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]
    assert run([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0]) == \
        [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]


def test_sampler_oracle_properties():
    n, n_items = 4000, 97
    a = philox_ref.sample_negatives(5, 0, n, 1, n_items)
    b = philox_ref.sample_negatives(5, 1, n, 1, n_items)
    assert a.min() >= 0 and a.max() < n_items and not np.array_equal(a, b)
    counts = np.bincount(a, minlength=n_items)
    assert counts.min() > 0.4 * n / n_items and counts.max() < 1.8 * n / n_items  # roughly uniform
    # rejection: every user rates all even items -> only odd items may come out
    users = np.arange(n) % 10
    pos_rp = np.arange(0, 11 * 49, 49)[:11].astype(np.int32)
    pos_ci = np.tile(np.arange(0, 97, 2)[:49], 10).astype(np.int32)
    c = philox_ref.sample_negatives(5, 0, n, 1, n_items, users, pos_rp, pos_ci, max_trials=64)
    assert (c % 2 == 1).all()
    unchanged = a % 2 == 1
    assert np.array_equal(c[unchanged], a[unchanged])  # accepted first candidates are untouched


def _dense(z, prefix):
    import scipy.sparse as sp
    shape = tuple(int(x) for x in z[f"{prefix}_shape"])
    return torch.from_numpy(sp.csr_matrix((z[f"{prefix}_data"], z[f"{prefix}_indices"], z[f"{prefix}_indptr"]), shape=shape).toarray()).double()


def test_mhcn_forward_matches_reference(golden):
    from oracle import social_ref

    z = golden("mhcn_model")
    names = [k[len("param__"):] for k in z.keys() if k.startswith("param__")]
    p = {n: T(z[f"param__{n}"].astype(np.float64), True) for n in names}
    perms = [torch.from_numpy(x) for x in z["perms"]]
    out = social_ref.mhcn_forward(p, _dense(z, "Hs"), _dense(z, "Hj"), _dense(z, "Hp"), _dense(z, "R"), int(z["n_layers"]),
                                  float(z["ss_rate"]), T(z["u_idx"]), T(z["v_idx"]), T(z["neg_idx"]), perms)
    for got, key in zip(out, ("batch_user", "batch_pos", "batch_neg", "ss_loss", "final_user", "final_item")):
        np.testing.assert_allclose(got.detach().numpy(), z[key], rtol=1e-4, atol=1e-5)
    total = losses_ref.bpr_log_eps_sigmoid(out[0], out[1], out[2]) + out[3]
    grads = torch.autograd.grad(total, [p[n] for n in names], allow_unused=True)
    for n, g in zip(names, grads):
        g = np.zeros_like(z[f"grad__{n}"]) if g is None else g.numpy()   # sgating_*.4 take no part in the forward
        np.testing.assert_allclose(g, z[f"grad__{n}"], rtol=2e-3, atol=2e-6)


def test_diffnet_forward_matches_reference(golden):
    from oracle import social_ref

    z = golden("diffnet_model")
    uw, iw = T(z["user_w"].astype(np.float64), True), T(z["item_w"].astype(np.float64), True)
    ws = [T(w.astype(np.float64), True) for w in z["weights"]]
    fu = social_ref.diffnet_forward(uw, iw, ws, _dense(z, "S"), _dense(z, "A"))
    np.testing.assert_allclose(fu.detach().numpy(), z["final_user"], rtol=1e-5, atol=1e-7)
    u, v, n = fu[T(z["u_idx"])], iw[T(z["i_idx"])], iw[T(z["j_idx"])]
    y = (u * v).sum(1) - (u * n).sum(1)
    loss = -torch.log(torch.sigmoid(y)).sum() + float(z["regU"]) * (u.norm(2) + v.norm(2) + n.norm(2))
    _check(loss, (uw, iw, *ws), z["loss"], (z["g_user_w"], z["g_item_w"], *z["g_weights"]), rtol=1e-4)


def test_eval_measures_match_reference(golden):
    from oracle import eval_ref

    z = golden("eval_metrics")
    ptr = z["test_ptr"]
    tests = [z["test_items"][ptr[u]:ptr[u + 1]] for u in range(len(ptr) - 1)]
    got = eval_ref.measures(z["lists"], tests, [int(n) for n in z["top_ns"]])
    for row, n in zip(z["measures"], z["top_ns"]):
        np.testing.assert_allclose(got[int(n)], row, rtol=0, atol=1e-5)


def test_buir_nb_matches_reference(golden):
    """BUIR_NB (univariate/buir.py:236-277) without dropout: online / target propagation, predictor, loss, momentum update."""
    z = golden("buir_nb")
    U, I, K, mom = int(z["n_users"]), int(z["n_items"]), int(z["n_layers"]), float(z["momentum"])
    rp, ci, v = z["norm_indptr"], z["norm_indices"], z["norm_data"]
    A = torch.zeros(U + I, U + I, dtype=torch.float64)
    A[torch.from_numpy(np.repeat(np.arange(U + I), np.diff(rp))), torch.from_numpy(ci.astype(np.int64))] = torch.from_numpy(v).double()

    def enc(uw, iw):
        x = torch.cat([uw, iw]); layers = [x]
        for _ in range(K):
            layers.append(A @ layers[-1])
        f = torch.stack(layers, 1).mean(1)
        return f[:U], f[U:]
    ou, oi = T(z["online_user"].astype(np.float64), True), T(z["online_item"].astype(np.float64), True)
    W, b = T(z["pred_w"].astype(np.float64), True), T(z["pred_b"].astype(np.float64), True)
    users, items = T(z["users"]), T(z["items"])
    uo, io = enc(ou, oi)
    ut, it = enc(T(z["target_user0"].astype(np.float64)), T(z["target_item0"].astype(np.float64)))
    out = (uo[users] @ W.T + b, ut[users], io[items] @ W.T + b, it[items])
    for got, key in zip(out, ("out_u_online", "out_u_target", "out_i_online", "out_i_target")):
        np.testing.assert_allclose(got.detach().numpy(), z[key], rtol=1e-4, atol=1e-5)
    n = lambda t: torch.nn.functional.normalize(t, dim=-1)
    loss = ((2 - 2 * (n(out[0]) * n(out[3])).sum(-1)) + (2 - 2 * (n(out[2]) * n(out[1])).sum(-1))).mean()
    _check(loss, (ou, oi, W, b), z["loss"], (z["g_user"], z["g_item"], z["g_pred_w"], z["g_pred_b"]), rtol=1e-4)
    t1 = z["target_user0"].copy(); t1[z["users"]] = z["target_user0"][z["users"]] * mom + z["online_user"][z["users"]] * (1 - mom)
    np.testing.assert_allclose(t1, z["target_user1"], rtol=1e-6, atol=1e-7)


def _golden_csr(z, prefix):
    import scipy.sparse as sp

    shape = tuple(int(x) for x in z[f"{prefix}_shape"])
    m = sp.csr_matrix((z[f"{prefix}_data"], z[f"{prefix}_indices"], z[f"{prefix}_indptr"]), shape=shape)
    m.sort_indices()
    return m


def test_motif_oracle_matches_reference_fixture(golden):
    """oracle/motif_ref.py against the reference's own MHCN.build_hyper_adj_mats (mhcn.py:340-368) run on the same S / Y
    (tests/golden/make_golden_motifs.py): identical sparsity pattern, values within 1 ulp."""
    from oracle import motif_ref

    z = golden("mhcn_motifs")
    S, Y = _golden_csr(z, "S"), _golden_csr(z, "Y")
    got = motif_ref.build_hyper_adj_mats(S, Y)
    for h, name in zip(got, ("Hs", "Hj", "Hp")):
        want = _golden_csr(z, name)
        want.eliminate_zeros()
        h.sort_indices()
        assert h.nnz == want.nnz > 0, name
        assert np.array_equal(h.indptr, want.indptr) and np.array_equal(h.indices, want.indices), name
        np.testing.assert_allclose(h.data, want.data, rtol=2e-7, atol=0)
    # the terms themselves: counts are non-negative integers, A1..A5, A8..A10 symmetric
    terms, B, U = motif_ref.motif_terms(S, Y)
    assert (B != B.T).nnz == 0 and B.multiply(U).nnz == 0
    for k, a in terms.items():
        assert np.all(a.data == np.round(a.data)) and np.all(a.data >= 0), k
        if k not in ("A6", "A7"):
            assert (a != a.T).nnz == 0, k


def test_sept_social_oracle_matches_reference(golden):
    """oracle/social_ref.py (sept_social part) against the reference's own SEPT class (tests/golden/make_golden_sept_social.py)."""
    import scipy.sparse as sp
    from oracle import social_ref

    z = golden("sept_social")
    bi, Y = _golden_csr(z, "bi"), _golden_csr(z, "Y")
    views = social_ref.sept_social_views(bi, Y)
    for got, name in zip(views, ("social", "sharing")):
        want = _golden_csr(z, name)
        got.sort_indices()
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices), name
        np.testing.assert_allclose(got.data, want.data, rtol=3e-7)
    U, I = int(z["user_num"]), int(z["item_num"])
    adj = torch.from_numpy(sp.coo_matrix((z["adj_data"], (z["adj_row"], z["adj_col"])), shape=(U + I, U + I)).toarray()).double()
    dense = lambda name: torch.from_numpy(_golden_csr(z, name).toarray()).double()
    uw = torch.from_numpy(z["user_w"]).double().requires_grad_(True)
    iw = torch.from_numpy(z["item_w"]).double().requires_grad_(True)
    t = lambda k: torch.from_numpy(z[k])
    out = social_ref.sept_social_iteration(uw, iw, adj, dense("social"), dense("sharing"), int(z["n_layers"]), float(z["ss_rate"]),
                                           int(z["ins_cnt"]), float(z["reg"]), t("user_idx"), t("pos_idx"), t("neg_idx"))
    for k in ("rec_user", "rec_item", "sharing_view", "friend_view", "social_prediction", "sharing_prediction", "rec_prediction"):
        np.testing.assert_allclose(out[k].detach().numpy(), z[k], rtol=2e-4, atol=2e-6, err_msg=k)
    for k, mk in (("f_pos", "f_margin"), ("sh_pos", "sh_margin"), ("r_pos", "r_margin")):
        clear = z[mk] > 1e-6                                # rows whose K-th / (K+1)-th probabilities are not fp32-close
        assert clear.mean() > 0.9
        assert np.array_equal(np.sort(out[k].numpy()[clear], 1), np.sort(z[k][clear], 1)), k
    for k in ("rec_loss", "nd_f", "nd_s", "nd_r", "total"):
        np.testing.assert_allclose(float(out[k]), float(z[k]), rtol=2e-4, err_msg=k)
    out["total"].backward()
    np.testing.assert_allclose(uw.grad.numpy(), z["g_user"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(iw.grad.numpy(), z["g_item"], rtol=2e-3, atol=2e-6)


def test_esrf_oracle_matches_reference(golden):
    """oracle restatements of esrf.py (motif adjacency, generator, discriminator, losses) against the reference's own
    functions / modules (tests/golden/make_golden_esrf.py)."""
    from oracle import motif_ref, social_ref

    z = golden("esrf")
    S, Y = _golden_csr(z, "S"), _golden_csr(z, "Y")
    A = motif_ref.build_motif_induced_adjacency_matrix(S, Y)
    want = _golden_csr(z, "A"); want.eliminate_zeros()
    A.sort_indices()
    assert np.array_equal(A.indptr, want.indptr) and np.array_equal(A.indices, want.indices)
    np.testing.assert_allclose(A.data, want.data, rtol=3e-7)
    K, seg, regU, beta = int(z["K"]), int(z["segment"]), float(z["regU"]), float(z["beta"])
    dn = lambda name: torch.from_numpy(_golden_csr(z, name).toarray()).double()
    t = lambda k: torch.from_numpy(z[k])
    rel = t("gen_relation").double().requires_grad_(True); sel = t("gen_selector").double().requires_grad_(True)
    uw = t("dis_user").double().requires_grad_(True); iw = t("dis_item").double().requires_grad_(True)
    alt = social_ref.esrf_generator(rel, sel, dn("A"), int(z["n_layers_G"]), seg, t("noise").double())
    np.testing.assert_allclose(alt.detach().numpy(), z["alt"], rtol=2e-3, atol=1e-6)
    pu, pi = social_ref.esrf_discriminator(uw, iw, dn("joint"), None, int(z["n_layers_D"]), False, K)
    np.testing.assert_allclose(pu.detach().numpy(), z["pre_user"], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(pi.detach().numpy(), z["pre_item"], rtol=1e-4, atol=1e-6)
    pair, reg, _, _ = social_ref.esrf_losses(pu, pi, alt.detach(), t("user_idx"), t("i_idx"), t("j_idx"), K, regU)
    np.testing.assert_allclose([float(pair), float(reg)], [float(z["pre_pair"]), float(z["pre_reg"])], rtol=1e-4)
    g = torch.autograd.grad(pair + reg, (uw, iw))
    np.testing.assert_allclose(g[0].numpy(), z["g_pre_user"], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(g[1].numpy(), z["g_pre_item"], rtol=2e-3, atol=1e-6)
    su, si = social_ref.esrf_discriminator(uw, iw, dn("joint"), alt, int(z["n_layers_D"]), True, K)
    np.testing.assert_allclose(su.detach().numpy(), z["soc_user"], rtol=1e-4, atol=1e-6)
    pair, reg, adv, g_adv = social_ref.esrf_losses(su, si, alt, t("user_idx"), t("i_idx"), t("j_idx"), K, regU)
    np.testing.assert_allclose([float(pair), float(reg), float(adv), float(beta * g_adv)],
                               [float(z["pair"]), float(z["reg"]), float(z["adv"]), float(z["g_loss"])], rtol=1e-4)
    gd = torch.autograd.grad(pair + reg + beta * adv, (uw, iw), retain_graph=True)
    np.testing.assert_allclose(gd[0].numpy(), z["g_d_user"], rtol=2e-3, atol=2e-6)
    np.testing.assert_allclose(gd[1].numpy(), z["g_d_item"], rtol=2e-3, atol=2e-6)
    gg = torch.autograd.grad(beta * g_adv, (rel, sel))
    np.testing.assert_allclose(gg[0].numpy(), z["g_g_relation"], rtol=5e-3, atol=1e-6 * np.abs(z["g_g_relation"]).max())
    np.testing.assert_allclose(gg[1].numpy(), z["g_g_selector"], rtol=5e-3, atol=1e-6 * np.abs(z["g_g_selector"]).max())


def test_ingest_oracle_matches_reference_loaders(golden):
    """oracle/ingest_ref.py against the reference's own load_data + Interaction (ncl.py: sorted string ids; selfcf.py: first
    appearance) on the same file bytes; and the packed-key order equals Python's string order."""
    from oracle import ingest_ref

    z = golden("ingest")
    train = ingest_ref.load_pairs(z["train_bytes"].tobytes())
    test = ingest_ref.load_pairs(z["test_bytes"].tobytes())
    assert len(train) == int(z["n_records"])
    for order in ("sorted", "appearance"):
        umap = ingest_ref.number_ids([u for u, _ in train], order)
        imap = ingest_ref.number_ids([i for _, i in train], order)
        assert [s for s, _ in sorted(umap.items(), key=lambda kv: kv[1])] == list(z[f"{order}_user_ids"])
        assert [s for s, _ in sorted(imap.items(), key=lambda kv: kv[1])] == list(z[f"{order}_item_ids"])
        assert np.array_equal([umap[u] for u, _ in train], z[f"{order}_users"]) and np.array_equal([imap[i] for _, i in train], z[f"{order}_items"])
        assert np.array_equal([umap.get(u, -1) for u, _ in test], z[f"{order}_test_users"])
        assert np.array_equal([imap.get(i, -1) for _, i in test], z[f"{order}_test_items"])
    ids = sorted({u for u, _ in train})
    assert sorted(ids, key=ingest_ref.pack_key) == ids
    with pytest.raises(ValueError):
        ingest_ref.pack_key("123456789")


def test_packed_key_round_trip_on_the_host():
    """ingest.decode_keys (host side of the GPU ingest path) inverts the oracle's key packing, also for keys whose top bit is set
    when read as signed 64-bit integers."""
    from oracle import ingest_ref
    from recommendation_b200 import ingest

    ids = ["1", "10", "2", "u9", "u10", "it07", "abcdefgh", "~zz"]
    keys = np.array([ingest_ref.pack_key(s) for s in ids], dtype=np.uint64)
    t = torch.from_numpy(keys.view(np.int64).copy())
    assert ingest.decode_keys(t) == ids
    order = np.argsort(keys, kind="stable")
    assert [ids[k] for k in order] == sorted(ids)               # unsigned key order == Python's string order


def test_key_tuples_of_long_ids_round_trip_and_order_on_the_host():
    """Ids longer than 8 bytes travel as tuples of 64-bit words (gcf_text_parse_pairs_words): tuple order == string order,
    and ingest.decode_keys inverts the packing for word-major [W, n] tensors."""
    from oracle import ingest_ref
    from recommendation_b200 import ingest

    ids = ["user_1", "user_10", "user_2", "customer-id:000123", "customer-id:00012", "A" * 17, "A" * 17 + "B", "b", "~" * 24]
    w = 3
    tuples = [ingest_ref.pack_words(s, w) for s in ids]
    assert [s for _, s in sorted(zip(tuples, ids))] == sorted(ids)
    arr = np.array(tuples, dtype=np.uint64).T.copy()              # [W, n] word-major
    assert ingest.decode_keys(torch.from_numpy(arr.view(np.int64))) == ids
    with pytest.raises(ValueError):
        ingest_ref.pack_words("x" * 25, 3)
