"""GPU parity: the CUDA path (through the C-ABI of libgcf.so) against the CPU oracle and the golden fixtures
generated from the reference.  Bit-exact for index / integer work; fp32 tolerances of the north star
(rtol 1e-3) for propagated embeddings, losses and gradients -- most checks are far tighter."""
import numpy as np
import pytest
import torch

from oracle import graph_ref, losses_ref, lightgcn_ref, philox_ref
from recommendation_b200 import _lib, functional as F_, synth
from recommendation_b200.graph import CSRGraph
from recommendation_b200.lightgcn import LightGCN, FusedLightGCNTrainer, build_edge_index, bpr_step_loss, train_step

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-3, 1e-6  # north-star tolerance for fp32 results
ULP2 = 2.4e-7


def dev_t(a, cuda, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(cuda)


def random_coo(rng, n_rows, n_cols, nnz, dup_frac=0.1, hub=False):
    r = rng.integers(0, n_rows, nnz)
    c = rng.integers(0, n_cols, nnz)
    if hub:  # a few very long rows / hot columns
        r[: nnz // 3] = rng.integers(0, 3, nnz // 3)
        c[nnz // 3: nnz // 2] = rng.integers(0, 2, nnz // 2 - nnz // 3)
    nd = int(nnz * dup_frac)
    if nd:
        src = rng.integers(0, nnz, nd)
        r = np.concatenate([r, r[src]]); c = np.concatenate([c, c[src]])
    v = (rng.random(r.shape[0]) + 0.5).astype(np.float32)
    return r.astype(np.int64), c.astype(np.int64), v


# ====================================================================== graph build
@pytest.mark.parametrize("with_vals", [False, True])
@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 0), (50, 70, 300), (1000, 900, 20000), (70000, 65000, 300000)])
def test_coo_to_csr_bit_exact(cuda, shape, with_vals):
    n_rows, n_cols, nnz = shape
    rng = np.random.default_rng(nnz + 1)
    r, c, v = random_coo(rng, n_rows, n_cols, nnz, hub=nnz > 1000)
    g = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda) if with_vals else None, n_rows, n_cols, norm="none")
    rp, ci, vals = graph_ref.coo_to_canonical_csr(r, c, v if with_vals else None, n_rows, n_cols)
    assert np.array_equal(g.row_ptr.cpu().numpy(), rp)
    assert np.array_equal(g.col_idx.cpu().numpy(), ci)
    # duplicates are summed in their original order on both sides -> bit-exact values
    assert np.array_equal(g.vals.cpu().numpy(), vals)


def test_degree_count_and_edge_index_bit_exact(cuda):
    inter = synth.power_law_bipartite(300, 500, 4000, seed=3)
    U, I = inter.n_users, inter.n_items
    ei = build_edge_index(dev_t(inter.users, cuda), dev_t(inter.items, cuda), U)
    want = graph_ref.bipartite_edge_index(inter.users, inter.items, U)
    assert np.array_equal(ei.cpu().numpy(), want)
    lib = _lib.load()
    deg = torch.empty(U + I, dtype=torch.int32, device=cuda)
    _lib.check(lib.gcf_degree_count(_lib.ptr(ei[1].contiguous()), ei.shape[1], _lib.ptr(deg), U + I, _lib.current_stream()), "deg")
    assert np.array_equal(deg.cpu().numpy(), graph_ref.degrees(want[1], U + I))


def test_sym_graph_matches_reference_fixture(cuda, golden):
    g = golden("selfcf_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    csr = CSRGraph.from_pairs(dev_t(g["u_idx"], cuda), dev_t(g["i_idx"], cuda), U, I, norm="sym")
    assert np.array_equal(csr.row_ptr.cpu().numpy(), g["norm_indptr"])
    assert np.array_equal(csr.col_idx.cpu().numpy(), g["norm_indices"])
    assert np.array_equal(csr.degrees().cpu().numpy(), g["rowsum"])          # degrees: bit-exact
    np.testing.assert_allclose(csr.vals.cpu().numpy(), g["norm_data"], rtol=ULP2, atol=0)  # <= 2 ulp


def test_raw_graph_matches_ncl_fixture(cuda, golden):
    g = golden("ncl_graph")
    U, I = int(g["n_users"]), int(g["n_items"])
    csr = CSRGraph.from_coo(dev_t(g["coo_row"], cuda), dev_t(g["coo_col"], cuda), dev_t(g["coo_data"], cuda), U + I, U + I, norm="none")
    rp, ci, v = graph_ref.coo_to_canonical_csr(g["coo_row"], g["coo_col"], g["coo_data"], U + I, U + I)
    assert np.array_equal(csr.row_ptr.cpu().numpy(), rp) and np.array_equal(csr.col_idx.cpu().numpy(), ci)
    assert np.array_equal(csr.vals.cpu().numpy(), v)


@pytest.mark.parametrize("mode", ["sym", "row"])
def test_norm_values_with_isolated_nodes(cuda, mode):
    rng = np.random.default_rng(5)
    n = 400
    r, c, v = random_coo(rng, n, n, 3000)
    r[r % 7 == 0] = 1  # rows 0, 7, 14, ... become empty -> rowsum 0 -> inf -> 0
    csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda), n, n, norm=mode)
    rp, ci, vals = graph_ref.coo_to_canonical_csr(r, c, v, n, n)
    want, rowsum, dinv = graph_ref.normalize_csr(rp, ci, vals, n, n, mode)
    assert (rowsum == 0).any()
    got = csr.vals.cpu().numpy()
    assert np.isfinite(got).all()
    np.testing.assert_allclose(got, want, rtol=2 * ULP2, atol=0)
    np.testing.assert_allclose(csr.rowsum.cpu().numpy(), rowsum, rtol=1e-6)


@pytest.mark.parametrize("shape", [(60, 90, 500), (5000, 3000, 100000)])
def test_csr_transpose_bit_exact(cuda, shape):
    n_rows, n_cols, nnz = shape
    rng = np.random.default_rng(11)
    r, c, v = random_coo(rng, n_rows, n_cols, nnz, hub=True)
    csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda), n_rows, n_cols, norm="none")
    t = csr.transpose()
    rp, ci, vals = graph_ref.coo_to_canonical_csr(r, c, v, n_rows, n_cols)
    trp, tci, tv = graph_ref.csr_transpose(rp, ci, vals, n_rows, n_cols)
    assert np.array_equal(t.row_ptr.cpu().numpy(), trp)
    assert np.array_equal(t.col_idx.cpu().numpy(), tci)
    assert np.array_equal(t.vals.cpu().numpy(), tv)


# ====================================================================== SpMM
@pytest.mark.parametrize("d", [16, 32, 64, 128, 256, 8, 48, 200, 320, 520])
def test_spmm_matches_oracle(cuda, d):
    rng = np.random.default_rng(d)
    n_rows, n_cols, nnz = 700, 500, 9000
    r, c, v = random_coo(rng, n_rows, n_cols, nnz, hub=True)
    csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda), n_rows, n_cols, norm="none", chunk=64)
    assert csr.plan.n_long > 0
    x = rng.standard_normal((n_cols, d)).astype(np.float32)
    y = torch.empty(n_rows, d, device=cuda)
    F_.spmm_raw(csr, dev_t(x, cuda), y=y)
    want = graph_ref.spmm_csr(csr.row_ptr.cpu().numpy(), csr.col_idx.cpu().numpy(), csr.vals.cpu().numpy(), x)
    np.testing.assert_allclose(y.cpu().numpy(), want, rtol=1e-4, atol=1e-4)
    # run twice: the self-resetting chunk tickets must leave the workspace reusable
    y2 = torch.empty_like(y)
    F_.spmm_raw(csr, dev_t(x, cuda), y=y2)
    assert torch.equal(y, y2), "long-row reduction must be deterministic and re-runnable"


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
@pytest.mark.parametrize("d", [8, 16, 32, 64, 128])
def test_spmm_variants_agree(cuda, d, variant):
    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=1)
    # chunk=64 puts the hub rows on the long-row path too (narrow-row variants batch several (col, val) pairs per lane)
    csr = CSRGraph.from_pairs(dev_t(inter.users, cuda), dev_t(inter.items, cuda), 3000, 4000, norm="sym",
                              chunk=64 if d <= 32 else None)
    x = torch.randn(7000, d, device=cuda)
    y0 = torch.empty_like(x); y1 = torch.empty_like(x)
    F_.spmm_raw(csr, x, y=y0, variant=0)
    F_.spmm_raw(csr, x, y=y1, variant=variant)
    assert torch.equal(y0, y1)  # same summation order in every variant


@pytest.mark.parametrize("shape", [(1000, 8, 8), (777, 4, 16), (5, 2, 32), (300, 1, 64), (64, 8, 4)])
def test_slices_rows_layout_conversion(cuda, shape):
    """gcf_slices_to_rows / gcf_rows_to_slices: [G][n][w] column slices <-> row-major [n, G*w] (pure data movement: bit-exact)."""
    n, G, w = shape
    lib, st = _lib.load(), _lib.current_stream()
    blocked = torch.randn(G, n, w, device=cuda)
    rows = torch.full((n, G * w + 4), -7.0, device=cuda)            # leading dimension larger than the row
    _lib.check(lib.gcf_slices_to_rows(_lib.ptr(blocked), _lib.ptr(rows), rows.stride(0), n, G, w, st), "gcf_slices_to_rows")
    want = blocked.permute(1, 0, 2).reshape(n, G * w)
    assert torch.equal(rows[:, : G * w], want) and bool((rows[:, G * w:] == -7.0).all())
    back = torch.empty_like(blocked)
    _lib.check(lib.gcf_rows_to_slices(_lib.ptr(rows), rows.stride(0), _lib.ptr(back), n, G, w, st), "gcf_rows_to_slices")
    assert torch.equal(back, blocked)
    assert lib.gcf_slices_to_rows(_lib.ptr(blocked), _lib.ptr(rows), 4, n, G, w, st) != 0   # ld too small is rejected


@pytest.mark.parametrize("shape", [(1000, 8, 8), (777, 4, 16), (5, 2, 32), (301, 3, 12), (64, 16, 4)])
def test_peer_movers_on_local_sources(cuda, shape):
    """csrc/peer.cu's three movers with every "peer" on this GPU (the multi-GPU trainers hand them IPC mappings of other ranks'
    buffers; the arithmetic is the same): gather_cols = slices -> rows, sum_cols = fixed-order sum of column slices,
    copy_blocks = row blocks of different heights stacked.  Pure data movement / one rounding per add: bit-exact."""
    import ctypes

    n, G, w = shape
    lib, st = _lib.load(), _lib.current_stream()
    # gather_cols: G sources [n, w] (leading dimension w + 4) -> [n, G*w]
    srcs = [torch.randn(n, w + 4, device=cuda) for _ in range(G)]
    rows = torch.full((n, G * w + 8), -7.0, device=cuda)
    _lib.check(lib.gcf_peer_gather_cols(_lib.ptr_array(srcs), G, w + 4, _lib.ptr(rows), rows.stride(0), n, w, st), "gcf_peer_gather_cols")
    want = torch.cat([t[:, :w] for t in srcs], dim=1)
    assert torch.equal(rows[:, : G * w], want) and bool((rows[:, G * w:] == -7.0).all())
    # sum_cols: column slice [lo, lo + w) of G full-width partials, summed in source order
    d = G * w
    parts = [torch.randn(n, d, device=cuda) for _ in range(G)]
    lo = (G - 1) * w
    ptrs = (ctypes.c_void_p * G)(*[t.data_ptr() + 4 * lo for t in parts])
    out = torch.full((n, w + 4), 3.0, device=cuda)
    _lib.check(lib.gcf_peer_sum_cols(ptrs, G, d, _lib.ptr(out), out.stride(0), n, w, st), "gcf_peer_sum_cols")
    acc = torch.zeros(n, w, device=cuda)
    for t in parts:
        acc = acc + t[:, lo:lo + w]
    assert torch.equal(out[:, :w], acc) and bool((out[:, w:] == 3.0).all())
    # copy_blocks: blocks of 0 .. n rows stacked in source order
    heights = [(n * (g + 1)) // (G + 1) if g % 3 != 1 else 0 for g in range(G)]
    ptrs = (ctypes.c_void_p * G)(*[t.data_ptr() + 4 * lo for t in parts])
    hts = (ctypes.c_int64 * G)(*heights)
    stack = torch.full((sum(heights) + 1, w), -1.0, device=cuda)
    _lib.check(lib.gcf_peer_copy_blocks(ptrs, hts, None, G, d, _lib.ptr(stack), w, w, st), "gcf_peer_copy_blocks")
    want = torch.cat([t[:h, lo:lo + w] for t, h in zip(parts, heights)], dim=0)
    assert torch.equal(stack[: sum(heights)], want) and bool((stack[sum(heights):] == -1.0).all())
    # the same with explicit destination offsets: source g fills rows g, g + G, ... (users dealt out cyclically)
    hts = (ctypes.c_int64 * G)(*[(n - g + G - 1) // G for g in range(G)])
    offs = (ctypes.c_int64 * G)(*[g * w for g in range(G)])
    inter = torch.full((n, w), -1.0, device=cuda)
    _lib.check(lib.gcf_peer_copy_blocks(ptrs, hts, offs, G, d, _lib.ptr(inter), G * w, w, st), "gcf_peer_copy_blocks")
    for g in range(G):
        assert torch.equal(inter[g::G], parts[g][: (n - g + G - 1) // G, lo:lo + w])
    assert lib.gcf_peer_gather_cols(_lib.ptr_array(srcs), G, w + 4, _lib.ptr(rows), w, n, w, st) != 0   # ld_dst too small is rejected


def test_peer_alloc_export_roundtrip(cuda):
    """gcf_peer_alloc hands out a zero-filled block torch can alias, gcf_peer_export a 64-byte IPC handle for it (opening it
    takes a second process: tests/test_dist.py)."""
    import ctypes

    from recommendation_b200 import peer

    blk = peer._Block(1000)
    t = torch.as_tensor(blk, device=cuda)
    assert t.data_ptr() == blk.ptr and t.numel() == 1000 and float(t.abs().sum()) == 0.0
    t.fill_(2.0)
    handle = ctypes.create_string_buffer(peer.HANDLE_BYTES)
    _lib.check(_lib.load().gcf_peer_export(ctypes.c_void_p(blk.ptr), handle), "gcf_peer_export")
    assert any(handle.raw)
    del t
    blk.free()


def test_spmm_epilogues(cuda):
    rng = np.random.default_rng(2)
    n, d = 600, 64
    r, c, v = random_coo(rng, n, n, 8000, hub=True)
    csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda), n, n, norm="row", chunk=128)
    x = torch.randn(n, d, device=cuda)
    z1, z2 = torch.randn(n, d, device=cuda), torch.randn(n, d, device=cuda)
    y = torch.empty(n, d, device=cuda); o = torch.empty(n, d, device=cuda)
    F_.spmm_raw(csr, x, y=y, out=o, epilogue=_lib.EPILOGUE_L2NORM, alpha=0.7, post=0.5, addends=[z1, z2], betas=[1.0, -2.0])
    t = graph_ref.spmm_csr(csr.row_ptr.cpu().numpy(), csr.col_idx.cpu().numpy(), csr.vals.cpu().numpy(), x.cpu().numpy())
    np.testing.assert_allclose(y.cpu().numpy(), t, rtol=1e-4, atol=1e-5)
    nrm = np.maximum(np.linalg.norm(t, axis=1, keepdims=True), 1e-12)
    want = 0.5 * (0.7 * t / nrm + z1.cpu().numpy().astype(np.float64) - 2.0 * z2.cpu().numpy())
    np.testing.assert_allclose(o.cpu().numpy(), want, rtol=1e-4, atol=1e-5)


def test_spmm_autograd_nonsymmetric(cuda):
    rng = np.random.default_rng(4)
    r, c, v = random_coo(rng, 300, 200, 4000, hub=True)
    csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), dev_t(v, cuda), 300, 200, norm="row")
    x = torch.randn(200, 32, device=cuda, requires_grad=True)
    w = torch.randn(300, 32, device=cuda)
    (F_.spmm(csr, x) * w).sum().backward()
    dense = torch.from_numpy(csr.to_scipy().toarray()).double()
    want = dense.T @ w.cpu().double()
    np.testing.assert_allclose(x.grad.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-5)


def test_empty_and_ragged_operators(cuda):
    # no entries at all, a single row, rows with exactly LPR / LPR+1 entries
    csr = CSRGraph.from_coo(torch.zeros(0, dtype=torch.int64, device=cuda), torch.zeros(0, dtype=torch.int64, device=cuda), None, 5, 5)
    x = torch.randn(5, 64, device=cuda)
    y = torch.full((5, 64), 7.0, device=cuda)
    F_.spmm_raw(csr, x, y=y)
    assert torch.count_nonzero(y) == 0
    for deg in (1, 15, 16, 17, 31, 32, 33, 255, 256, 257):
        r = np.zeros(deg, np.int64); c = np.arange(deg, dtype=np.int64)
        csr = CSRGraph.from_coo(dev_t(r, cuda), dev_t(c, cuda), None, 2, 300)
        x = torch.randn(300, 64, device=cuda)
        y = torch.empty(2, 64, device=cuda)
        F_.spmm_raw(csr, x, y=y)
        np.testing.assert_allclose(y[0].cpu().numpy(), x[:deg].double().sum(0).cpu().numpy(), rtol=1e-4, atol=1e-5)
        assert torch.count_nonzero(y[1]) == 0


# ====================================================================== propagation (fixtures from the reference)
def test_propagate_matches_selfcf_encoder_fixture(cuda, golden):
    g, e = golden("selfcf_graph"), golden("selfcf_encoder")
    U, I = int(g["n_users"]), int(g["n_items"])
    csr = CSRGraph.from_pairs(dev_t(g["u_idx"], cuda), dev_t(g["i_idx"], cuda), U, I, norm="sym")
    uw = dev_t(e["user_w"], cuda).requires_grad_(True)
    iw = dev_t(e["item_w"], cuda).requires_grad_(True)
    final = F_.propagate(csr, torch.cat([uw, iw]), int(e["n_layers"]), mode="mean")
    np.testing.assert_allclose(final[:U].detach().cpu().numpy(), e["user_all"], rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(final[U:].detach().cpu().numpy(), e["item_all"], rtol=RTOL, atol=ATOL)
    ((final[:U] * dev_t(e["proj_u"], cuda)).sum() + (final[U:] * dev_t(e["proj_i"], cuda)).sum()).backward()
    np.testing.assert_allclose(uw.grad.cpu().numpy(), e["grad_user_w"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(iw.grad.cpu().numpy(), e["grad_item_w"], rtol=RTOL, atol=1e-5)


def test_propagate_matches_ncl_encoder_fixture(cuda, golden):
    g, e = golden("ncl_graph"), golden("ncl_encoder")
    U, I = int(g["n_users"]), int(g["n_items"])
    csr = CSRGraph.from_coo(dev_t(g["coo_row"], cuda), dev_t(g["coo_col"], cuda), dev_t(g["coo_data"], cuda), U + I, U + I,
                            norm="none", symmetric=True)
    uw = dev_t(e["user_w"], cuda).requires_grad_(True)
    iw = dev_t(e["item_w"], cuda).requires_grad_(True)
    x0 = torch.cat([uw, iw])
    final, layers = F_.propagate(csr, x0, int(e["n_layers"]), mode="mean", return_layers=True)
    all_emb = [x0] + layers
    np.testing.assert_allclose(final[:U].detach().cpu().numpy(), e["user_out"], rtol=RTOL, atol=1e-5)
    np.testing.assert_allclose(final[U:].detach().cpu().numpy(), e["item_out"], rtol=RTOL, atol=1e-5)
    for k, t in enumerate(all_emb):
        np.testing.assert_allclose(t.detach().cpu().numpy(), e["all_emb"][k], rtol=RTOL, atol=1e-5)
    proj = dev_t(e["proj_layers"], cuda)
    loss = (final[:U] * dev_t(e["proj_u"], cuda)).sum() + (final[U:] * dev_t(e["proj_i"], cuda)).sum()
    loss = loss + sum((t * proj[k]).sum() for k, t in enumerate(all_emb))
    loss.backward()  # gradients reach every layer output: exercises the `extra` path of gcf_propagate_bwd
    np.testing.assert_allclose(uw.grad.cpu().numpy(), e["grad_user_w"], rtol=RTOL, atol=1e-4)
    np.testing.assert_allclose(iw.grad.cpu().numpy(), e["grad_item_w"], rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("k,mode", [(1, "mean"), (2, "sum"), (3, "sum"), (4, "mean"), (6, "mean")])
def test_propagate_vs_oracle_small_graph(cuda, k, mode):
    inter = synth.power_law_bipartite(3000, 4000, 100000, seed=k)
    U, I, d = 3000, 4000, 64
    csr = CSRGraph.from_pairs(dev_t(inter.users, cuda), dev_t(inter.items, cuda), U, I, norm="sym")
    x0 = (torch.randn(U + I, d) * 0.1)
    xg = x0.to(cuda).requires_grad_(True)
    final = F_.propagate(csr, xg, k, mode=mode)
    rp, ci, v = csr.row_ptr.cpu().numpy(), csr.col_idx.cpu().numpy(), csr.vals.cpu().numpy()
    _, want = graph_ref.propagate(rp, ci, v, x0.numpy(), k, mode)
    np.testing.assert_allclose(final.detach().cpu().numpy(), want, rtol=RTOL, atol=1e-6)
    w = torch.randn(U + I, d)
    (final * w.to(cuda)).sum().backward()
    _, gwant = graph_ref.propagate(rp, ci, v, w.numpy(), k, mode)  # symmetric operator
    np.testing.assert_allclose(xg.grad.cpu().numpy(), gwant, rtol=RTOL, atol=1e-5)


# ====================================================================== gather / scatter / sampler / adam
@pytest.mark.parametrize("d", [16, 64, 128, 200])
@pytest.mark.parametrize("det", [False, True])
def test_gather_and_scatter_add(cuda, d, det):
    rng = np.random.default_rng(d)
    table = torch.randn(500, d, device=cuda, requires_grad=True)
    idx_np = np.concatenate([rng.integers(0, 500, 3000), np.full(400, 7), np.full(300, 499)])  # hot rows
    rng.shuffle(idx_np)
    idx = dev_t(idx_np, cuda)
    out = F_.gather_rows(table, idx)
    assert torch.equal(out, table.detach()[idx])  # gather: bit-exact
    w = torch.randn(idx.numel(), d, device=cuda)
    F_.set_deterministic(det)
    try:
        (out * w).sum().backward()
    finally:
        F_.set_deterministic(False)
    want = torch.zeros(500, d, dtype=torch.float64).index_add_(0, torch.from_numpy(idx_np), w.cpu().double())
    np.testing.assert_allclose(table.grad.cpu().numpy(), want.numpy(), rtol=1e-4, atol=1e-4)
    if det:
        table2 = table.detach().clone().requires_grad_(True)
        F_.set_deterministic(True)
        try:
            (F_.gather_rows(table2, idx) * w).sum().backward()
        finally:
            F_.set_deterministic(False)
        assert torch.equal(table.grad, table2.grad), "deterministic scatter-add must be run-to-run identical"


def test_gather_accepts_python_lists(cuda):
    table = torch.randn(50, 32, device=cuda)
    assert torch.equal(F_.gather_rows(table, [3, 3, 49, 0]), table[[3, 3, 49, 0]])
    assert F_.gather_rows(table, []).shape == (0, 32)


def test_sampler_bit_exact_vs_oracle(cuda):
    n, n_items = 5000, 4099
    got = F_.sample_negatives(n, n_items, seed=0x1234567890ABCDEF, offset=3, device=cuda)
    want = philox_ref.sample_negatives(0x1234567890ABCDEF, 3, n, 1, n_items)
    assert np.array_equal(got.cpu().numpy(), want)
    got = F_.sample_negatives(n, n_items, seed=9, offset=(1 << 40) + 5, n_negs=3, device=cuda)
    want = philox_ref.sample_negatives(9, (1 << 40) + 5, n, 3, n_items).reshape(n, 3)
    assert np.array_equal(got.cpu().numpy(), want)
    # rejection against per-user sorted positives
    inter = synth.power_law_bipartite(200, 300, 6000, seed=8)
    rp, ci, _ = graph_ref.coo_to_canonical_csr(inter.users, inter.items, None, 200, 300)
    users = np.random.default_rng(0).integers(0, 200, n)
    got = F_.sample_negatives(n, 300, seed=77, offset=1, users=dev_t(users, cuda), positives=(dev_t(rp, cuda), dev_t(ci, cuda)),
                              max_trials=100, device=cuda)
    want = philox_ref.sample_negatives(77, 1, n, 1, 300, users, rp, ci, max_trials=100)
    assert np.array_equal(got.cpu().numpy(), want)
    pos = set(zip(inter.users.tolist(), inter.items.tolist()))
    assert not any((u, j) in pos for u, j in zip(users.tolist(), got.cpu().tolist()))


def test_sampler_window_equals_slice_of_the_full_draw(cuda):
    """gcf_sample_negatives_at: a rank that owns triples [t0, t1) draws exactly that window of the single-GPU stream."""
    from recommendation_b200 import _lib
    lib = _lib.load()
    n, n_items, n_negs = 10_000, 4321, 3
    full = F_.sample_negatives(n, n_items, seed=99, offset=7, n_negs=n_negs, device=cuda)
    for t0, t1 in ((0, 17), (17, 5000), (5000, n)):
        out = torch.empty((t1 - t0) * n_negs, dtype=torch.int64, device=cuda)
        _lib.check(lib.gcf_sample_negatives_at(99, 7, t0 * n_negs, None, t1 - t0, n_negs, n_items, None, None, 1, _lib.ptr(out),
                                               _lib.current_stream()), "gcf_sample_negatives_at")
        assert torch.equal(out.view(-1, n_negs), full[t0:t1])
    # the oracle's stream definition, at a window that crosses 2^32 slots
    base = (1 << 32) - 5
    out = torch.empty(10, dtype=torch.int64, device=cuda)
    _lib.check(lib.gcf_sample_negatives_at(3, 1, base, None, 10, 1, 1000, None, None, 1, _lib.ptr(out), _lib.current_stream()),
               "gcf_sample_negatives_at")
    from oracle import philox_ref
    want = philox_ref.sample_negatives(3, 1, 10, 1, 1000, slot_base=base)
    assert out.cpu().tolist() == want.tolist()
    # gcf_sample_negatives_pos: an interleaved subset of the triples (users dealt out cyclically over the ranks)
    for r, G in ((0, 2), (3, 8)):
        pos = torch.arange(r, n, G, dtype=torch.int64, device=cuda)
        out = torch.empty(pos.numel() * n_negs, dtype=torch.int64, device=cuda)
        _lib.check(lib.gcf_sample_negatives_pos(99, 7, _lib.ptr(pos), pos.numel(), n_negs, n_items, _lib.ptr(out),
                                                _lib.current_stream()), "gcf_sample_negatives_pos")
        assert torch.equal(out.view(-1, n_negs), full[pos])


def test_xavier_table_is_a_function_of_seed_row_and_column(cuda):
    """sharded trainers must not repeat rows across shards; any column slice equals the full table's."""
    from recommendation_b200.tables import xavier_uniform_table
    U, I, d = 5000, 3000, 64
    full = xavier_uniform_table(U, I, d, seed=5, device=cuda, chunk_rows=1024)
    for lo, hi in ((0, 8), (8, 16), (32, 64)):
        assert torch.equal(xavier_uniform_table(U, I, d, seed=5, device=cuda, cols=(lo, hi), chunk_rows=1024), full[:, lo:hi])
    assert torch.unique(full, dim=0).shape[0] == U + I                                  # no duplicate rows
    bu, bi = (6.0 / (U + d)) ** 0.5, (6.0 / (I + d)) ** 0.5
    assert float(full[:U].abs().max()) <= bu and float(full[U:].abs().max()) <= bi
    assert float(full[:U].abs().max()) > 0.99 * bu and float(full[U:].abs().max()) > 0.99 * bi
    assert abs(float(full[:U].std()) / (bu / 3 ** 0.5) - 1) < 0.02                          # uniform(-b, b): std = b / sqrt(3)
    assert not torch.equal(full, xavier_uniform_table(U, I, d, seed=6, device=cuda, chunk_rows=1024))


@pytest.mark.parametrize("wd,decoupled", [(0.0, False), (1e-2, False), (1e-2, True)])
def test_adam_matches_torch(cuda, wd, decoupled):
    torch.manual_seed(0)
    p0 = torch.randn(1000, 64)
    ref = p0.clone().requires_grad_(True)
    opt = (torch.optim.AdamW if decoupled else torch.optim.Adam)([ref], lr=0.01, weight_decay=wd)
    p = p0.to(cuda); m = torch.zeros_like(p); v = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(1000, 64)
        ref.grad = g.clone(); opt.step()
        F_.adam_step_(p, g.to(cuda), m, v, step, lr=0.01, weight_decay=wd, decoupled=decoupled)
    np.testing.assert_allclose(p.cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("momentum,wd,nesterov,dampening", [(0.9, 0.0, False, 0.0), (0.9, 1e-3, False, 0.0), (0.9, 1e-3, True, 0.0),
                                                              (0.5, 0.0, False, 0.1), (0.0, 1e-2, False, 0.0)])
def test_sgd_momentum_matches_torch(cuda, momentum, wd, nesterov, dampening):
    """torch.optim.SGD(lr, momentum=0.9) of selfcf.py:544 / directau.py:214, through the drop-in optimiser class."""
    from recommendation_b200 import optim
    torch.manual_seed(1)
    p0 = torch.randn(777, 33)                      # odd size: exercises the scalar tail of the vector loop
    ref = p0.clone().requires_grad_(True)
    kw = dict(lr=0.05, momentum=momentum, weight_decay=wd, nesterov=nesterov, dampening=dampening)
    opt_ref = torch.optim.SGD([ref], **kw)
    p = torch.nn.Parameter(p0.to(cuda))
    opt = optim.SGD([p], **kw)
    for _ in range(6):
        g = torch.randn(777, 33)
        ref.grad = g.clone(); opt_ref.step()
        p.grad = g.to(cuda); opt.step()
    np.testing.assert_allclose(p.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    if momentum:
        np.testing.assert_allclose(opt.state[p]["momentum_buffer"].cpu().numpy(),
                                   opt_ref.state[ref]["momentum_buffer"].numpy(), rtol=1e-5, atol=1e-6)


def test_adam_optimizer_class_matches_torch(cuda):
    from recommendation_b200 import optim
    torch.manual_seed(2)
    p0 = torch.randn(513, 64)
    ref = p0.clone().requires_grad_(True)
    opt_ref = torch.optim.Adam([ref], lr=0.01)
    p = torch.nn.Parameter(p0.to(cuda))
    opt = optim.Adam([p], lr=0.01)
    for _ in range(4):
        g = torch.randn(513, 64)
        ref.grad = g.clone(); opt_ref.step()
        p.grad = g.to(cuda); opt.step()
    np.testing.assert_allclose(p.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    sd = opt.state_dict()["state"][0]
    assert set(sd) == {"step", "exp_avg", "exp_avg_sq"} and float(sd["step"]) == 4.0


def test_row_sparse_adam_touches_only_listed_rows(cuda):
    """gcf_adam_rows_step == torch.optim.Adam's arithmetic on the listed rows; every other row (and its moments) untouched.
    (torch.optim.SparseAdam itself uses another denominator, sqrt(v) + eps without the bias correction, so the reference
    arithmetic is dense Adam run on the sub-table of the listed rows.)"""
    torch.manual_seed(3)
    n, d = 1000, 64
    p0 = torch.randn(n, d)
    rows = torch.randperm(n)[:137]
    p = p0.to(cuda); m = torch.zeros_like(p); v = torch.zeros_like(p)
    ref = p0[rows].clone().requires_grad_(True)
    opt_ref = torch.optim.Adam([ref], lr=0.01, weight_decay=1e-3)
    for step in range(1, 5):
        g_rows = torch.randn(rows.numel(), d)
        ref.grad = g_rows.clone(); opt_ref.step()
        F_.adam_rows_step_(p, rows.to(cuda), g_rows.to(cuda), m, v, step, lr=0.01, weight_decay=1e-3)
    got = p.cpu()
    np.testing.assert_allclose(got[rows].numpy(), ref.detach().numpy(), rtol=1e-5, atol=1e-6)
    untouched = torch.ones(n, dtype=torch.bool); untouched[rows] = False
    assert torch.equal(got[untouched], p0[untouched])
    assert float(m.cpu()[untouched].abs().max()) == 0.0 and float(v.cpu()[untouched].abs().max()) == 0.0


# ====================================================================== BPR
def test_bpr_matches_ncl_fixture(cuda, golden):
    z = golden("ncl_losses")
    ue, pe, ne = (dev_t(z[k], cuda).requires_grad_(True) for k in ("ue", "pe", "ne"))
    loss = F_.bpr_loss_rows(ue, pe, ne, variant="log_eps_sigmoid", eps=1e-5)
    np.testing.assert_allclose(loss.item(), z["bpr"], rtol=RTOL)
    loss.backward()
    for t, k in ((ue, "g_bpr_u"), (pe, "g_bpr_p"), (ne, "g_bpr_n")):
        np.testing.assert_allclose(t.grad.cpu().numpy(), z[k], rtol=RTOL, atol=1e-7)


def test_bpr_matches_gcl_fixture(cuda, golden):
    z = golden("gcl_losses")
    ue, pe, ne = (dev_t(z[k], cuda).requires_grad_(True) for k in ("ue", "pe", "ne"))
    b = ue.shape[0]
    items = torch.cat([pe, ne])
    ar = torch.arange(b, device=cuda)
    r = float(z["reg_weight"]) / b
    loss = F_.bpr_loss_gather(ue, items, ar, ar, ar + b, variant="softplus", reg_u=r, reg_p=r, reg_n=r)
    np.testing.assert_allclose(loss.item(), z["bpr_reg"], rtol=RTOL)
    loss.backward()
    for t, k in ((ue, "g_u"), (pe, "g_p"), (ne, "g_n")):
        np.testing.assert_allclose(t.grad.cpu().numpy(), z[k], rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("d", [16, 64, 128, 72])
@pytest.mark.parametrize("n_negs", [1, 3])
@pytest.mark.parametrize("sorted_users", [False, True])
def test_bpr_lightgcn_form_vs_oracle(cuda, d, n_negs, sorted_users):
    rng = np.random.default_rng(d + n_negs)
    U, I, T = 300, 400, 5003
    ue = torch.randn(U, d) * 0.3; ie = torch.randn(I, d) * 0.3
    pu = rng.integers(0, U, T); pi = rng.integers(0, I, T)
    if sorted_users:
        pu = np.sort(pu)
    ni = rng.integers(0, I, (T, n_negs)) if n_negs > 1 else rng.integers(0, I, T)
    ue_c, ie_c = ue.clone().requires_grad_(True), ie.clone().requires_grad_(True)
    want = losses_ref.bpr_lightgcn(ue_c.double(), ie_c.double(), torch.from_numpy(pu), torch.from_numpy(pi), torch.from_numpy(ni), 1e-3)
    want.backward()
    ue_g, ie_g = ue.to(cuda).requires_grad_(True), ie.to(cuda).requires_grad_(True)
    got = bpr_step_loss(ue_g, ie_g, dev_t(pu, cuda), dev_t(pi, cuda), dev_t(ni, cuda), 1e-3)
    np.testing.assert_allclose(got.item(), want.item(), rtol=1e-5)
    got.backward()
    np.testing.assert_allclose(ue_g.grad.cpu().numpy(), ue_c.grad.numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ie_g.grad.cpu().numpy(), ie_c.grad.numpy(), rtol=RTOL, atol=1e-7)


def test_bpr_empty_batch(cuda):
    ue = torch.randn(10, 64, device=cuda, requires_grad=True); ie = torch.randn(10, 64, device=cuda, requires_grad=True)
    e = torch.zeros(0, dtype=torch.int64, device=cuda)
    assert F_.bpr_loss_gather(ue, ie, e, e, e).item() == 0.0


# ====================================================================== LightGCN model + step
def _tiny_problem(seed=0, U=300, I=500, E=4000):
    inter = synth.power_law_bipartite(U, I, E, seed=seed)
    return inter, torch.from_numpy(inter.users), torch.from_numpy(inter.items)


def test_lightgcn_forward_backward_vs_oracle(cuda):
    inter, pu, pi = _tiny_problem()
    U, I, d, K = inter.n_users, inter.n_items, 64, 3
    torch.manual_seed(1)
    model = LightGCN(U, I, d, K).to(cuda)
    assert sorted(model.state_dict().keys()) == ["item_embedding.weight", "user_embedding.weight"]
    uw = model.user_embedding.weight.detach().cpu().clone().requires_grad_(True)
    iw = model.item_embedding.weight.detach().cpu().clone().requires_grad_(True)
    ei = build_edge_index(pu, pi, U)
    neg = torch.from_numpy(np.random.default_rng(0).integers(0, I, pu.numel()))
    want = lightgcn_ref.lightgcn_step_loss(uw, iw, ei, pu, pi, neg, K, 1e-4)
    want.backward()
    ue, ie = model(ei.to(cuda))
    ue_w, ie_w = lightgcn_ref.lightgcn_forward(uw.detach(), iw.detach(), ei, K)
    np.testing.assert_allclose(ue.detach().cpu().numpy(), ue_w.numpy(), rtol=RTOL, atol=1e-6)
    np.testing.assert_allclose(ie.detach().cpu().numpy(), ie_w.numpy(), rtol=RTOL, atol=1e-6)
    got = bpr_step_loss(ue, ie, pu.to(cuda), pi.to(cuda), neg.to(cuda), 1e-4)
    np.testing.assert_allclose(got.item(), want.item(), rtol=1e-5)
    got.backward()
    np.testing.assert_allclose(model.user_embedding.weight.grad.cpu().numpy(), uw.grad.numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(model.item_embedding.weight.grad.cpu().numpy(), iw.grad.numpy(), rtol=RTOL, atol=1e-7)


def test_train_steps_track_reference_optimisation(cuda):
    """Three full optimiser steps (forward, BPR+reg, backward, Adam) with identical pre-drawn negatives:
    public-API path, fused-trainer path and the CPU restatement of lightgcn.py stay together."""
    inter, pu, pi = _tiny_problem(seed=2)
    U, I, d, K = inter.n_users, inter.n_items, 64, 3
    torch.manual_seed(3)
    model = LightGCN(U, I, d, K).to(cuda)
    uw = model.user_embedding.weight.detach().cpu().clone().requires_grad_(True)
    iw = model.item_embedding.weight.detach().cpu().clone().requires_grad_(True)
    ref_opt = torch.optim.Adam([uw, iw], lr=0.01)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    ei = build_edge_index(pu, pi, U)
    ei_c = ei.to(cuda)
    graph = model.graph_for(ei_c)
    fused = FusedLightGCNTrainer(graph, U, I, model.table.detach().clone(), pu, pi, n_layers=K, lr=0.01, reg_weight=1e-4)
    cfg = {"n_neg": 1, "reg_weight": 1e-4, "loss_type": "bpr"}
    rng = np.random.default_rng(5)
    for step in range(3):
        neg = torch.from_numpy(rng.integers(0, I, pu.numel()))
        ref_opt.zero_grad()
        want = lightgcn_ref.lightgcn_step_loss(uw, iw, ei, pu, pi, neg, K, 1e-4)
        want.backward(); ref_opt.step()
        got = train_step(model, opt, ei_c, pu.to(cuda), pi.to(cuda), I, cfg, neg_i=neg.to(cuda))
        got_f = fused.step(neg_i=neg.to(cuda))
        np.testing.assert_allclose(got.item(), want.item(), rtol=1e-4)
        np.testing.assert_allclose(got_f.item(), want.item(), rtol=1e-4)
    np.testing.assert_allclose(model.user_embedding.weight.detach().cpu().numpy(), uw.detach().numpy(), rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(model.item_embedding.weight.detach().cpu().numpy(), iw.detach().numpy(), rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(fused.table[:U].cpu().numpy(), uw.detach().numpy(), rtol=RTOL, atol=2e-5)
    np.testing.assert_allclose(fused.table[U:].cpu().numpy(), iw.detach().numpy(), rtol=RTOL, atol=2e-5)


# ====================================================================== full-size properties (BASELINE cfg 1)
def test_cfg1_size_properties(cuda):
    """Size-independent properties at the full Gowalla-shaped size (the oracle comparison at this size is
    tests/test_gpu_fullsize_oracle.py): structural identities of the CSR, linearity and symmetry of the operator, and a
    float64 spot check of rows."""
    inter, d, K = synth.config_graph("cfg1")
    U, I = inter.n_users, inter.n_items
    csr = CSRGraph.from_pairs(dev_t(inter.users, cuda), dev_t(inter.items, cuda), U, I, norm="sym")
    assert csr.nnz == 2 * inter.n_edges  # synthetic pairs are unique
    rp = csr.row_ptr.cpu().numpy(); ci = csr.col_idx.cpu().numpy()
    deg = np.concatenate([np.bincount(inter.users, minlength=U), np.bincount(inter.items, minlength=I)])
    assert np.array_equal(np.diff(rp), deg)                      # degrees bit-exact
    assert np.array_equal(csr.degrees().cpu().numpy(), deg.astype(np.float32))
    row_of = np.repeat(np.arange(U + I), deg)
    assert (np.diff(ci)[np.diff(row_of) == 0] > 0).all()          # strictly ascending columns inside every row
    t = csr.transpose()
    assert t is csr
    x = torch.randn(U + I, d, device=cuda); y = torch.randn(U + I, d, device=cuda)
    ax, ay, axy = (torch.empty_like(x) for _ in range(3))
    F_.spmm_raw(csr, x, y=ax); F_.spmm_raw(csr, y, y=ay); F_.spmm_raw(csr, x + 2 * y, y=axy)
    torch.testing.assert_close(axy, ax + 2 * ay, rtol=1e-4, atol=1e-5)                      # linearity
    torch.testing.assert_close((y * ax).sum().double(), (x * ay).sum().double(), rtol=1e-4, atol=1e-3)  # <y,Ax> = <Ay,x>
    rows = np.random.default_rng(0).integers(0, U + I, 200)
    rows = np.concatenate([rows, np.argsort(-deg)[:5]])          # include the hub rows (long path)
    v = csr.vals.cpu().numpy(); xc = x.cpu().numpy().astype(np.float64)
    for r in rows:
        want = (v[rp[r]:rp[r + 1], None].astype(np.float64) * xc[ci[rp[r]:rp[r + 1]]]).sum(0)
        np.testing.assert_allclose(ax[r].cpu().numpy(), want, rtol=1e-4, atol=1e-5)


# ====================================================================== fused BPR forward+backward
def test_fused_bpr_equals_separate_forward_backward(cuda):
    rng = np.random.default_rng(11)
    U, I, T, d = 500, 700, 20011, 64
    lib = _lib.load(); st = _lib.current_stream()
    ue = torch.randn(U, d, device=cuda) * 0.3; ie = torch.randn(I, d, device=cuda) * 0.3
    pu = dev_t(np.sort(rng.integers(0, U, T)), cuda); pi = dev_t(rng.integers(0, I, T), cuda); ni = dev_t(rng.integers(0, I, T), cuda)
    ws_bytes = lib.gcf_bpr_workspace_bytes(T); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    loss_a = torch.empty((), device=cuda); coef = torch.empty(T, device=cuda)
    gu_a, gi_a = torch.zeros_like(ue), torch.zeros_like(ie)
    _lib.check(lib.gcf_bpr_fwd(_lib.ptr(ue), d, _lib.ptr(ie), d, d, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1, 1, 0.0, 0,
                               1e-3, 1e-3, 2e-3, _lib.ptr(loss_a), _lib.ptr(coef), _lib.ptr(ws), ws_bytes, st), "fwd")
    g = torch.tensor(0.5, device=cuda)
    _lib.check(lib.gcf_bpr_bwd(_lib.ptr(ue), d, _lib.ptr(ie), d, d, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1, _lib.ptr(coef),
                               _lib.ptr(g), 1e-3, 1e-3, 2e-3, _lib.ptr(gu_a), d, _lib.ptr(gi_a), d, st), "bwd")
    loss_b = torch.empty((), device=cuda); coef_b = torch.empty(T, device=cuda)
    gu_b, gi_b = torch.zeros_like(ue), torch.zeros_like(ie)
    _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(ue), d, _lib.ptr(ie), d, d, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1, 1, 0.0, 0,
                                   1e-3, 1e-3, 2e-3, 0.5, _lib.ptr(loss_b), _lib.ptr(coef_b), _lib.ptr(gu_b), d, _lib.ptr(gi_b), d,
                                   _lib.ptr(ws), ws_bytes, st), "fwd_bwd")
    np.testing.assert_allclose(loss_b.item(), loss_a.item(), rtol=1e-6)
    torch.testing.assert_close(coef_b, coef, rtol=1e-6, atol=1e-12)
    torch.testing.assert_close(gu_b, gu_a, rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(gi_b, gi_a, rtol=1e-4, atol=1e-8)


def test_fused_bpr_applies_upstream_gradient(cuda):
    """(3 * loss).backward(): the gradients computed in the forward launch are scaled by the upstream scalar."""
    rng = np.random.default_rng(12)
    U, I, T, d = 200, 300, 4001, 32
    ue = torch.randn(U, d) * 0.3; ie = torch.randn(I, d) * 0.3
    pu = rng.integers(0, U, T); pi = rng.integers(0, I, T); ni = rng.integers(0, I, T)
    ue_c, ie_c = ue.clone().requires_grad_(True), ie.clone().requires_grad_(True)
    (3.0 * losses_ref.bpr_lightgcn(ue_c.double(), ie_c.double(), torch.from_numpy(pu), torch.from_numpy(pi), torch.from_numpy(ni), 1e-3)).backward()
    ue_g, ie_g = ue.to(cuda).requires_grad_(True), ie.to(cuda).requires_grad_(True)
    loss = bpr_step_loss(ue_g, ie_g, dev_t(pu, cuda), dev_t(pi, cuda), dev_t(ni, cuda), 1e-3)
    (3.0 * loss).backward()
    np.testing.assert_allclose(ue_g.grad.cpu().numpy(), ue_c.grad.numpy(), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(ie_g.grad.cpu().numpy(), ie_c.grad.numpy(), rtol=RTOL, atol=1e-7)
    with torch.no_grad():   # no gradient requested: plain forward kernel, same value
        loss2 = bpr_step_loss(ue_g.detach(), ie_g.detach(), dev_t(pu, cuda), dev_t(pi, cuda), dev_t(ni, cuda), 1e-3)
    np.testing.assert_allclose(loss2.item(), loss.item(), rtol=1e-6)


# ====================================================================== feature-sharded building blocks on ONE GPU
@pytest.mark.parametrize("n_shards", [2, 4, 8])
def test_feature_slices_reproduce_the_full_step(cuda, n_shards):
    """The pieces FeatureShardedLightGCNTrainer runs per rank (d/G-wide SpMM, raw partial scores, loss from the summed
    scores, d/G-wide gradient pass), executed slice after slice on one GPU, against the unsharded kernels."""
    inter = synth.power_law_bipartite(400, 600, 9000, seed=8)
    U, I, d, K = 400, 600, 64, 2
    dg = d // n_shards
    lib = _lib.load(); st = _lib.current_stream()
    g = CSRGraph.from_pairs(dev_t(inter.users, cuda), dev_t(inter.items, cuda), U, I, norm="sym", chunk=64)
    x = torch.randn(U + I, d, device=cuda) * 0.3
    full = F_.propagate(g, x, K, mode="sum")
    rng = np.random.default_rng(3)
    T = inter.n_edges
    order = np.argsort(inter.users, kind="stable")
    pu, pi = dev_t(inter.users[order], cuda), dev_t(inter.items[order], cuda)
    ni = dev_t(rng.integers(0, I, T), cuda)
    reg = 1e-3
    ws_bytes = lib.gcf_bpr_workspace_bytes(T); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=cuda)
    # unsharded reference: fused forward+backward on the full-width tables
    want_loss = torch.empty((), device=cuda); want_g = torch.zeros_like(full)
    _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(full[:U]), d, _lib.ptr(full[U:]), d, d, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1,
                                   _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_MEAN, reg, reg, 0.0, 1.0, _lib.ptr(want_loss), None,
                                   _lib.ptr(want_g[:U]), d, _lib.ptr(want_g[U:]), d, _lib.ptr(ws), ws_bytes, st), "fused")
    scores = torch.zeros(T, device=cuda); reg_sum = torch.zeros((), device=cuda)
    finals = []
    for r in range(n_shards):
        xs = x[:, r * dg:(r + 1) * dg].contiguous()
        fs = F_.propagate(g, xs, K, mode="sum")                       # no exchange between slices
        torch.testing.assert_close(fs, full[:, r * dg:(r + 1) * dg], rtol=1e-5, atol=1e-6)
        finals.append(fs)
        part = torch.empty(T, device=cuda); lr = torch.empty((), device=cuda)
        _lib.check(lib.gcf_bpr_fwd(_lib.ptr(fs[:U]), dg, _lib.ptr(fs[U:]), dg, dg, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1,
                                   _lib.BPR_RAW_SCORE, 0.0, _lib.REDUCE_SUM, reg, reg, 0.0, _lib.ptr(lr), _lib.ptr(part),
                                   _lib.ptr(ws), ws_bytes, st), "raw scores")
        scores += part; reg_sum += lr                                   # = the all-reduce
    coef = torch.empty(T, device=cuda); loss_pt = torch.empty((), device=cuda)
    _lib.check(lib.gcf_bpr_coef_from_scores(_lib.ptr(scores), T, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_MEAN, _lib.ptr(loss_pt),
                                            _lib.ptr(coef), _lib.ptr(ws), ws_bytes, st), "coef")
    np.testing.assert_allclose((loss_pt + reg_sum).item(), want_loss.item(), rtol=1e-5)
    for r, fs in enumerate(finals):
        gs = torch.zeros_like(fs)
        _lib.check(lib.gcf_bpr_bwd(_lib.ptr(fs[:U]), dg, _lib.ptr(fs[U:]), dg, dg, _lib.ptr(pu), _lib.ptr(pi), _lib.ptr(ni), T, 1,
                                   _lib.ptr(coef), None, reg, reg, 0.0, _lib.ptr(gs[:U]), dg, _lib.ptr(gs[U:]), dg, st), "bwd")
        torch.testing.assert_close(gs, want_g[:, r * dg:(r + 1) * dg], rtol=1e-3, atol=1e-8)


def test_raw_score_mode_rejects_mean_reduction(cuda):
    lib = _lib.load()
    t = torch.zeros(4, 8, device=cuda); i = torch.zeros(2, dtype=torch.int64, device=cuda); o = torch.zeros(2, device=cuda)
    ws = torch.empty(lib.gcf_bpr_workspace_bytes(2), dtype=torch.uint8, device=cuda)
    rc = lib.gcf_bpr_fwd(_lib.ptr(t), 8, _lib.ptr(t), 8, 8, _lib.ptr(i), _lib.ptr(i), _lib.ptr(i), 2, 1, _lib.BPR_RAW_SCORE, 0.0,
                         _lib.REDUCE_MEAN, 0.0, 0.0, 0.0, _lib.ptr(o), _lib.ptr(o), _lib.ptr(ws), ws.numel(), _lib.current_stream())
    assert rc == -1 and b"raw scores need reduction = sum" in lib.gcf_last_error()


def test_fused_adam_epilogue_equals_separate_adam(cuda):
    """gcf_propagate_bwd_adam (Adam applied inside the last backward SpMM) against gcf_propagate_bwd + gcf_adam_step."""
    inter, pu, pi = _tiny_problem(seed=5, U=400, I=600, E=9000)
    U, I, d, K = inter.n_users, inter.n_items, 64, 3
    g = CSRGraph.from_pairs(pu.to(cuda), pi.to(cuda), U, I, norm="sym", chunk=64)      # chunk=64: long rows take the chunked path
    assert g.plan.n_long > 0
    t0 = torch.randn(U + I, d, device=cuda) * 0.05
    a = FusedLightGCNTrainer(g, U, I, t0.clone(), pu, pi, n_layers=K, lr=0.01, reg_weight=1e-4, weight_decay=1e-3, fused_adam=True)
    b = FusedLightGCNTrainer(g, U, I, t0.clone(), pu, pi, n_layers=K, lr=0.01, reg_weight=1e-4, weight_decay=1e-3, fused_adam=False)
    rng = np.random.default_rng(9)
    for s in range(3):
        neg = torch.from_numpy(rng.integers(0, I, pu.numel())).to(cuda)
        la, lb = a.step(neg_i=neg), b.step(neg_i=neg)
        np.testing.assert_allclose(la.item(), lb.item(), rtol=1e-4)
        if s == 0:
            # same gradient values feed the same update function; only the atomics of the BPR scatter reorder fp32 sums
            for x, y in ((a.exp_avg, b.exp_avg), (a.exp_avg_sq, b.exp_avg_sq)):
                assert ((x - y).norm() / y.norm()).item() < 1e-5
    err = (a.table - b.table).abs()
    frac = (err <= 2e-5 + 1e-3 * b.table.abs()).float().mean().item()
    assert frac > 0.999 and err.max().item() <= 0.07, f"fraction within tolerance {frac}, max deviation {err.max().item()}"


# ====================================================================== error behaviour of the C-ABI
def test_c_abi_reports_errors_instead_of_crashing(cuda):
    lib = _lib.load(); st = _lib.current_stream()
    inter = synth.power_law_bipartite(50, 60, 400, seed=1)
    g = CSRGraph.from_pairs(dev_t(inter.users, cuda), dev_t(inter.items, cuda), 50, 60, norm="sym", chunk=8)
    x = torch.randn(110, 6, device=cuda); y = torch.empty_like(x)
    rc = lib.gcf_spmm_csr_f32(g.struct_ref(), 6, _lib.ptr(x), 6, _lib.ptr(y), 6, None, 0, 0, 1.0, 1.0, 0, None, None, None, 0, 0, st)
    assert rc == -4 and b"d=6 unsupported" in lib.gcf_last_error()                       # GCF_EUNSUPPORTED
    x = torch.randn(110, 64, device=cuda); y = torch.empty_like(x)
    assert g.plan.n_long > 0
    rc = lib.gcf_spmm_csr_f32(g.struct_ref(), 64, _lib.ptr(x), 64, _lib.ptr(y), 64, None, 0, 0, 1.0, 1.0, 0, None, None, None, 0, 0, st)
    assert rc == -3 and b"workspace too small" in lib.gcf_last_error()                   # GCF_EWORKSPACE
    rc = lib.gcf_spmm_csr_f32(g.struct_ref(), 64, _lib.ptr(x), 64, None, 0, None, 0, 0, 1.0, 1.0, 0, None, None, None, 0, 0, st)
    assert rc == -1 and b"no output requested" in lib.gcf_last_error()                   # GCF_EINVAL
    with pytest.raises(_lib.GcfError, match="temperature must be positive"):
        F_.infonce_stats_raw(torch.randn(4, 16, device=cuda), torch.randn(4, 16, device=cuda), 0.0)
    with pytest.raises(_lib.GcfError, match="unsupported"):
        F_.infonce_stats_raw(torch.randn(4, 2048, device=cuda), torch.randn(4, 2048, device=cuda), 0.2)
    with pytest.raises(ValueError):
        F_.spmm(g, torch.randn(7, 64, device=cuda))                                       # row count mismatch
    with pytest.raises(RuntimeError, match="CUDA"):
        F_.gather_rows(torch.randn(4, 8), torch.tensor([0]))                             # CPU tensors are rejected
    # out-of-range COO entries raise (as scipy / torch indexing do in the reference) instead of corrupting memory or being
    # dropped silently
    r = torch.tensor([0, 1, 200, -1], device=cuda); c = torch.tensor([1, 0, 0, 0], device=cuda)
    with pytest.raises(ValueError, match="out of range"):
        CSRGraph.from_coo(r, c, None, 3, 3, norm="none")
    torch.cuda.synchronize()   # nothing above left a sticky CUDA error behind
    ok = CSRGraph.from_coo(r[:2], c[:2], None, 3, 3, norm="none")
    assert ok.nnz == 2 and torch.isfinite(F_.spmm(ok, torch.ones(3, 16, device=cuda))).all()


# ====================================================================== text ingest
def _ingest_files(tmp_path, z, prefix=""):
    tr, te = tmp_path / "train.txt", tmp_path / "test.txt"
    tr.write_bytes(z[f"{prefix}train_bytes"].tobytes()); te.write_bytes(z[f"{prefix}test_bytes"].tobytes())
    return str(tr), str(te)


@pytest.mark.parametrize("order", ["sorted", "appearance"])
def test_ingest_matches_reference_loaders(cuda, golden, tmp_path, order):
    """ingest.DeviceInteraction on the file bytes the reference's own load_data + Interaction read (ncl.py / selfcf.py):
    id dictionaries, dense training / test indices and the adjacency, all bit-exact."""
    import scipy.sparse as sp
    from recommendation_b200 import ingest

    z = golden("ingest")
    d = ingest.DeviceInteraction.from_files(*_ingest_files(tmp_path, z), id_order=order)
    assert d.user_ids() == list(z[f"{order}_user_ids"]) and d.item_ids() == list(z[f"{order}_item_ids"])
    assert d.user_num == len(z[f"{order}_user_ids"]) and d.item_num == len(z[f"{order}_item_ids"])
    for got, key in ((d.users, "users"), (d.items, "items"), (d.test_users, "test_users"), (d.test_items, "test_items")):
        assert np.array_equal(got.cpu().numpy(), z[f"{order}_{key}"]), key
    n = d.user_num + d.item_num
    if order == "sorted":        # raw adjacency of ncl.py:76-85: duplicates kept in the COO, summed by the CSR build
        want = sp.coo_matrix((z["sorted_adj_data"], (z["sorted_adj_row"], z["sorted_adj_col"])), shape=(n, n)).tocsr()
        got = d.norm_adj
    else:                        # selfcf.py: D^-1/2 (R + R^T) D^-1/2
        want = sp.csr_matrix((z["appearance_adj_data"], z["appearance_adj_indices"], z["appearance_adj_indptr"]), shape=(n, n))
        got = d.normalized_adj()
    want.sum_duplicates(); want.sort_indices()
    assert np.array_equal(got.row_ptr.cpu().numpy(), want.indptr) and np.array_equal(got.col_idx.cpu().numpy(), want.indices)
    np.testing.assert_allclose(got.vals.cpu().numpy(), want.data, rtol=3e-7)
    u, i, j = next(iter(d.sampler().batches(64)))              # the device sampler runs on the ingested pairs
    assert u.numel() == 64 and int(j.max()) < d.item_num


def test_ingest_numeric_ids_and_errors(cuda, golden, tmp_path):
    from recommendation_b200 import ingest

    z = golden("ingest")
    d = ingest.DeviceInteraction.from_files(*_ingest_files(tmp_path, z, "num_"), id_order="numeric")   # lightgcn.py:29-33
    assert np.array_equal(d.users.cpu().numpy(), z["num_users"]) and np.array_equal(d.items.cpu().numpy(), z["num_items"])
    assert (d.user_num, d.item_num) == (int(z["num_user_count"]), int(z["num_item_count"]))
    assert d.edge_index().shape == (2, 2 * d.users.numel())
    t = lambda s: torch.frombuffer(bytearray(s.encode()), dtype=torch.uint8).to(cuda)
    a, b = ingest.parse_pairs(t("user_123456 7 1\n"))          # an 11-byte id: two key words
    assert a.shape == (2, 1) and b.shape == (2, 1) and ingest.decode_keys(a) == ["user_123456"] and ingest.decode_keys(b) == ["7"]
    with pytest.raises(ValueError, match="MAX_ID_BYTES"):
        ingest.parse_pairs(t("x" * 300 + " 7 1\n"))
    with pytest.raises(ValueError, match="fewer than two"):
        ingest.parse_pairs(t("a b 1\nlonely\n"))
    with pytest.raises(ValueError, match="decimal"):
        ingest.parse_pairs(t("12 x7 1\n"), numeric=True)
    a, b = ingest.parse_pairs(torch.empty(0, dtype=torch.uint8, device=cuda))
    assert a.numel() == 0 and b.numel() == 0
    # a large random file against the oracle: record count, keys and numbering
    from oracle import ingest_ref
    rng = np.random.default_rng(9)
    rows = [f"{rng.integers(0, 50000)}\t{rng.integers(0, 9000)} 1" + ("\r" if k % 7 == 0 else "") for k in range(200000)]
    blob = ("\n".join(rows) + "\n").encode()
    di = ingest.DeviceInteraction(torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(cuda), id_order="sorted")
    pairs = ingest_ref.load_pairs(blob)
    um, im = ingest_ref.number_ids([p[0] for p in pairs], "sorted"), ingest_ref.number_ids([p[1] for p in pairs], "sorted")
    assert di.users.numel() == len(pairs) and di.user_num == len(um) and di.item_num == len(im)
    assert np.array_equal(di.users.cpu().numpy(), [um[p[0]] for p in pairs]) and np.array_equal(di.items.cpu().numpy(), [im[p[1]] for p in pairs])


@pytest.mark.parametrize("order", ["sorted", "appearance"])
def test_ingest_ids_longer_than_eight_bytes(cuda, order):
    """ncl.py:55-62 numbers arbitrary id STRINGS (sorted(set(ids))); selfcf.py:281-288 by first appearance.  Ids of 1..40
    bytes, shared prefixes, a test file with ids the training file does not hold (and longer ones) -- against the oracle."""
    from oracle import ingest_ref
    from recommendation_b200 import ingest

    rng = np.random.default_rng(21)
    stems = ["u", "user_", "customer-id:", "A" * 17, "A" * 17 + "B", "x" * 33 + "_"]
    users = [stems[rng.integers(0, len(stems))] + str(rng.integers(0, 3000)) for _ in range(40000)]
    items = ["item" + str(rng.integers(0, 800)).zfill(int(rng.integers(1, 12))) for _ in range(40000)]
    blob = "".join(f"{u} {i} 1\n" for u, i in zip(users, items)).encode()
    test_blob = (f"{users[5]} {items[7]} 1\nnever_seen_user_with_a_long_name {items[0]} 1\n{users[9]} {'z' * 50} 1\n").encode()
    t = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).to(cuda)
    di = ingest.DeviceInteraction(t(blob), t(test_blob), id_order=order)
    pairs = ingest_ref.load_pairs(blob)
    um, im = ingest_ref.number_ids([p[0] for p in pairs], order), ingest_ref.number_ids([p[1] for p in pairs], order)
    assert (di.user_num, di.item_num) == (len(um), len(im))
    assert np.array_equal(di.users.cpu().numpy(), [um[p[0]] for p in pairs])
    assert np.array_equal(di.items.cpu().numpy(), [im[p[1]] for p in pairs])
    assert di.test_users.cpu().tolist() == [um[users[5]], -1, um[users[9]]]
    assert di.test_items.cpu().tolist() == [im[items[7]], im[items[0]], -1]
    inv_u = {v: k for k, v in um.items()}
    assert di.user_ids() == [inv_u[k] for k in range(len(um))]                            # id2user, strings decoded from the key tuples
    assert di.norm_adj.nnz > 0
