"""Flat-stream SpMM (r02 default, csrc/spmm.cu spmm_flat_kernel) against the r01 row-walking kernel (variant 4) and a
dense fp64 product, on operators built to hit every branch of the tile walk: empty rows (compact row numbering),
hub rows (chunk schedule), rows of one entry, tiles of every size, widths that need the lane guard, rectangular
operators, and the three epilogue classes (plain store / linear combination with addends / row-L2-normalise).

Replaces torch.sparse.mm (ncl.py:419, selfcf.py:479, directau.py:290, mhcn.py:440-456)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

pytestmark = pytest.mark.gpu


def _random_operator(rng, n_rows, n_cols, avg_deg, n_hubs, hub_deg, p_empty):
    deg = rng.poisson(avg_deg, n_rows)
    deg[rng.random(n_rows) < p_empty] = 0
    if n_hubs:
        deg[rng.choice(n_rows, n_hubs, replace=False)] = rng.integers(hub_deg // 2, hub_deg + 1, n_hubs)
    deg = np.minimum(deg, n_cols)
    rows = np.repeat(np.arange(n_rows), deg)
    cols = np.concatenate([rng.choice(n_cols, k, replace=False) for k in deg]) if rows.size else np.zeros(0, np.int64)
    vals = rng.standard_normal(rows.size).astype(np.float32)
    return sp.csr_matrix((vals, (rows, cols)), shape=(n_rows, n_cols))


CASES = [
    # n_rows, n_cols, avg_deg, n_hubs, hub_deg, p_empty, chunk, tile_nnz
    (1, 5, 3, 0, 0, 0.0, 256, 64),
    (37, 41, 1, 0, 0, 0.5, 256, 16),
    (500, 300, 6, 3, 250, 0.0, 64, 32),
    (500, 300, 6, 3, 250, 0.2, 64, 32),
    (2000, 2000, 13, 10, 900, 0.05, 256, 96),
    (3000, 1500, 2, 0, 0, 0.3, 256, 512),
    (1200, 900, 30, 5, 600, 0.0, 128, 64),
    (64, 64, 0, 0, 0, 1.0, 256, 64),          # nothing but empty rows
    (300, 300, 1, 300, 40, 0.0, 16, 16),      # every row is a hub row: no tiles at all
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("d", [64, 128, 40, 100, 8, 16, 32])
def test_flat_kernel_matches_row_kernel_and_dense(cuda, case, d):
    from recommendation_b200 import _lib, functional as F_
    from recommendation_b200.graph import CSRGraph

    n_rows, n_cols, avg_deg, n_hubs, hub_deg, p_empty, chunk, tile_nnz = case
    rng = np.random.default_rng(hash(case) % 2**32 + d)
    mat = _random_operator(rng, n_rows, n_cols, avg_deg, n_hubs, hub_deg, p_empty)
    g = CSRGraph.from_scipy(mat, device=cuda, chunk=chunk, tile_nnz=tile_nnz)
    assert g.nnz == mat.nnz
    x = torch.from_numpy(rng.standard_normal((n_cols, d)).astype(np.float32)).to(cuda)
    adds = [torch.from_numpy(rng.standard_normal((n_rows, d)).astype(np.float32)).to(cuda) for _ in range(3)]
    dense = torch.from_numpy(mat.toarray().astype(np.float64)).to(cuda)
    want = dense @ x.double()

    def run(variant, **kw):
        y = torch.full((n_rows, d), float("nan"), device=cuda)
        o = torch.full((n_rows, d), float("nan"), device=cuda)
        F_.spmm_raw(g, x, y=y if kw.pop("want_y", True) else None, out=o if kw.pop("want_o", False) else None, variant=variant, **kw)
        return y, o

    scale = float(want.abs().max()) + 1.0
    # ask for the flat kernel explicitly (variants >= 10): the default dispatch sends rows of 8 / 16 floats, and launches with
    # an epilogue on L2-resident tables like these, to the row kernel (faster there)
    fv = 11 if d in (8, 16) else 12
    # same lanes per row in both kernels (d = 64, 128) -> same summation order -> bit-identical; the guarded widths use
    # 16 lanes per row in the flat kernel and 32 in the row kernel, so hub-row chunks are summed in another order
    same = (lambda a, b: torch.equal(a, b)) if d in (64, 128, 8, 16, 32) else (lambda a, b: torch.allclose(a, b, rtol=1e-5, atol=1e-5 * scale))
    # plain
    y_flat, _ = run(fv)
    y_row, _ = run(4)
    assert same(y_flat, y_row)
    assert float((y_flat.double() - want).abs().max()) <= 1e-5 * scale
    # Y and OUT with addends
    for n_add in (1, 3):
        kw = dict(want_o=True, alpha=1.5, post=0.5, addends=adds[:n_add], betas=[0.25, -1.0, 2.0][:n_add])
        yf, of = run(fv, **kw)
        yr, orow = run(4, **kw)
        assert same(yf, yr) and same(of, orow)
        ref = 0.5 * (1.5 * want + sum(b * a.double() for b, a in zip([0.25, -1.0, 2.0], adds[:n_add])))
        assert float((of.double() - ref).abs().max()) <= 1e-5 * (scale + 4.0)
    # OUT only, row-L2-normalised
    _, of = run(fv, want_y=False, want_o=True, epilogue=_lib.EPILOGUE_L2NORM)
    _, orow = run(4, want_y=False, want_o=True, epilogue=_lib.EPILOGUE_L2NORM)
    ref = want / want.norm(dim=1, keepdim=True).clamp_min(1e-12)
    assert torch.allclose(of, orow, rtol=0, atol=1e-6)
    assert float((of.double() - ref).abs().max()) <= 1e-5


def test_flat_kernel_duplicate_entries_and_row_strides(cuda):
    """ncl.py:76-85 keeps duplicate interactions (summed by the build); X / Y with a leading dimension > d."""
    from recommendation_b200 import functional as F_
    from recommendation_b200.graph import CSRGraph

    rng = np.random.default_rng(5)
    n, d, ld = 700, 64, 96
    rows = rng.integers(0, n, 9000); cols = (rng.zipf(1.5, 9000) - 1) % n
    mat = sp.coo_matrix((np.ones(9000, np.float32), (rows, cols)), shape=(n, n))
    g = CSRGraph.from_scipy(mat, device=cuda, chunk=64, tile_nnz=48)
    xbuf = torch.from_numpy(rng.standard_normal((n, ld)).astype(np.float32)).to(cuda)
    ybuf = torch.zeros(n, ld, device=cuda)
    F_.spmm_raw(g, xbuf[:, :d], y=ybuf[:, :d])
    want = torch.from_numpy(mat.toarray().astype(np.float64)).to(cuda) @ xbuf[:, :d].double()
    assert float((ybuf[:, :d].double() - want).abs().max()) <= 1e-5 * (float(want.abs().max()) + 1)
    assert float(ybuf[:, d:].abs().max()) == 0.0


@pytest.mark.parametrize("d", [64, 128])
def test_hub_flagged_gathers_are_bit_identical(cuda, d):
    """Hub flags only change the cache policy of the gathers (L1::evict_last for hub columns, L1::no_allocate for the rest):
    same columns, same order, same bits as the unflagged kernel (variant 15) and as the row kernel (variant 4)."""
    from recommendation_b200 import _lib, functional as F_
    from recommendation_b200.graph import CSRGraph, hub_flagged_columns

    rng = np.random.default_rng(7 + d)
    deg = rng.poisson(12.0, 3000); deg[rng.choice(3000, 6, replace=False)] = 900      # a few hub ROWS (chunked path) as well
    rows = np.repeat(np.arange(3000), deg)
    cols = (rng.zipf(1.3, rows.size) - 1) % 3000                                         # skewed column popularity
    mat = sp.coo_matrix((rng.standard_normal(rows.size).astype(np.float32), (rows, cols)), shape=(3000, 3000)).tocsr()
    mat.sum_duplicates()
    g = CSRGraph.from_scipy(mat, device=cuda, chunk=256, tile_nnz=96, hubs=64)
    assert g._hub_col_idx is not None
    flagged = g._hub_col_idx
    assert torch.equal(flagged & 0x7FFFFFFF, g.col_idx)                       # bit 31 is the only difference
    n_flag = int((flagged < 0).sum())
    assert 0 < n_flag < g.nnz
    # every flagged entry references one of the 64 most-referenced columns of its half
    deg = torch.bincount(g.col_idx.long(), minlength=3000)
    assert int(deg[g.col_idx[flagged < 0].long()].min()) >= 128
    plain = CSRGraph.from_scipy(mat, device=cuda, chunk=256, tile_nnz=96, hubs=0)
    assert plain._hub_col_idx is None
    x = torch.from_numpy(rng.standard_normal((3000, d)).astype(np.float32)).to(cuda)
    add = torch.from_numpy(rng.standard_normal((3000, d)).astype(np.float32)).to(cuda)
    outs = []
    for graph, variant in ((g, 0), (g, 16), (g, 15), (plain, 0), (g, 4)):
        y = torch.empty(3000, d, device=cuda); o = torch.empty(3000, d, device=cuda)
        F_.spmm_raw(graph, x, y=y, variant=variant)
        F_.spmm_raw(graph, x, out=o, alpha=0.5, addends=[add], betas=[2.0], variant=variant)
        outs.append((y, o))
    for y, o in outs[1:]:
        assert torch.equal(y, outs[0][0]) and torch.equal(o, outs[0][1])
    assert hub_flagged_columns(g.row_ptr, g.col_idx, 3000, 64, min_degree=10**9) is None
