"""Motif-induced adjacency matrices of MHCN on the GPU.

Drop-in for `MHCN.build_hyper_adj_mats` (univariate/mhcn.py:340-368): the reference evaluates sixteen
`(P.dot(Q)).multiply(M)` terms and the full `Y.dot(Y.T)` with scipy on the host; here every term is one
`gcf_spgemm_masked` launch on M's pattern (the product is never materialised), `S.multiply(S.T)` / `S - B` are
`gcf_csr_sample` look-ups, the full product is expand-sort-compress (`gcf_spgemm_count/_expand` + the stable COO -> CSR
build, which also performs every sparse sum below), and the row normalisation is `gcf_norm_values`.  torch only
concatenates / filters index arrays.  The result feeds `social.MHCNModel` directly.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .graph import CSRGraph

Coo = Tuple[torch.Tensor, torch.Tensor, torch.Tensor]   # (rows int64, cols int64, vals fp32)

# largest number of scalar products expanded at once (20 bytes each + the sort workspace)
MAX_PRODUCTS_PER_CHUNK = 1 << 27


def entry_rows(g: CSRGraph) -> torch.Tensor:
    """Row id of every stored entry (int64 [nnz])."""
    counts = (g.row_ptr[1:] - g.row_ptr[:-1]).to(torch.int64)
    return torch.repeat_interleave(torch.arange(g.n_rows, dtype=torch.int64, device=g.device), counts)


def to_coo(g: CSRGraph, vals: Optional[torch.Tensor] = None, *, drop_zeros: bool = True) -> Coo:
    v = g.vals if vals is None else vals
    r, c = entry_rows(g), g.col_idx.to(torch.int64)
    if drop_zeros:   # scipy's multiply / binary ops never store a zero result
        keep = v != 0
        r, c, v = r[keep], c[keep], v[keep]
    return r, c, v


def transpose_coo(t: Coo) -> Coo:
    return t[1], t[0], t[2]


def from_coo_sum(terms: Sequence[Coo], n_rows: int, n_cols: int, device, *, norm: str = "none") -> CSRGraph:
    """Sparse sum of COO terms as one canonical CSR (duplicates summed by gcf_coo_to_csr_stable)."""
    if terms:
        r = torch.cat([t[0] for t in terms]); c = torch.cat([t[1] for t in terms]); v = torch.cat([t[2] for t in terms])
    else:
        r = c = torch.empty(0, dtype=torch.int64, device=device); v = torch.empty(0, dtype=torch.float32, device=device)
    return CSRGraph.from_coo(r.contiguous(), c.contiguous(), v.contiguous(), n_rows, n_cols, norm=norm)


def csr_sample(x: CSRGraph, pattern: CSRGraph) -> torch.Tensor:
    """x[i, j] at every stored entry (i, j) of `pattern` (0 where x stores nothing)."""
    out = torch.zeros(max(pattern.nnz, 1), dtype=torch.float32, device=pattern.device)[: pattern.nnz]
    _lib.check(_lib.load().gcf_csr_sample(x.struct_ref(), pattern.struct_ref(), _lib.ptr(out) if pattern.nnz else None,
                                          _lib.current_stream()), "gcf_csr_sample")
    return out


def masked_product(a: CSRGraph, bt: CSRGraph, mask: CSRGraph) -> torch.Tensor:
    """Values of (A . B) o M on M's pattern; `bt` is the CSR of B^T.  (mhcn.py:345-360, one call per term.)"""
    out = torch.zeros(max(mask.nnz, 1), dtype=torch.float32, device=mask.device)[: mask.nnz]
    _lib.check(_lib.load().gcf_spgemm_masked(a.struct_ref(), bt.struct_ref(), mask.struct_ref(),
                                             _lib.ptr(out) if mask.nnz else None, _lib.current_stream()), "gcf_spgemm_masked")
    return out


def spgemm(a: CSRGraph, b: CSRGraph, *, max_products: int = MAX_PRODUCTS_PER_CHUNK) -> CSRGraph:
    """C = A . B (expand-sort-compress), row blocks of A sized to at most `max_products` scalar products each."""
    if a.n_cols != b.n_rows:
        raise ValueError(f"spgemm: inner dimensions differ ({a.n_cols} vs {b.n_rows})")
    if not 0 < max_products < 2 ** 32:
        raise ValueError("spgemm: max_products must be in (0, 2^32): product offsets inside a row block are 32-bit")
    lib, st, dev = _lib.load(), _lib.current_stream(), a.device
    row_ptr = a.row_ptr.cpu().tolist()
    n_prod = torch.zeros(1, dtype=torch.int64, device=dev)
    pieces: List[Coo] = []

    def run(r0: int, r1: int) -> None:
        e0, e1 = row_ptr[r0], row_ptr[r1]
        if e1 == e0:
            return
        ws_bytes = lib.gcf_spgemm_workspace_bytes(e1 - e0)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gcf_spgemm_count(a.struct_ref(), b.struct_ref(), e0, e1, _lib.ptr(n_prod), _lib.ptr(ws), ws_bytes, st),
                   "gcf_spgemm_count")
        total = int(n_prod.item())
        if total == 0:
            return
        if total > max_products:
            if r1 - r0 == 1:
                raise RuntimeError(f"spgemm: row {r0} alone needs {total} products (> max_products = {max_products})")
            mid = (r0 + r1) // 2
            del ws
            run(r0, mid); run(mid, r1)
            return
        rows = torch.empty(total, dtype=torch.int64, device=dev)
        cols = torch.empty(total, dtype=torch.int64, device=dev)
        vals = torch.empty(total, dtype=torch.float32, device=dev)
        _lib.check(lib.gcf_spgemm_expand(a.struct_ref(), b.struct_ref(), e0, e1, _lib.ptr(rows), _lib.ptr(cols), _lib.ptr(vals),
                                         _lib.ptr(ws), ws_bytes, st), "gcf_spgemm_expand")
        del ws
        block = CSRGraph.from_coo(rows, cols, vals, a.n_rows, b.n_cols, norm="none")   # sorts + sums this row block
        del rows, cols, vals
        pieces.append(to_coo(block, drop_zeros=False))

    run(0, a.n_rows)
    return from_coo_sum(pieces, a.n_rows, b.n_cols, dev)


class MotifTerms:
    """The terms every motif-based social model of the reference sums (mhcn.py:343-359 = esrf.py:1072-1088), as COO lists:
    social = the terms of A1..A7 (A1, A2, A3, A5 already symmetrised), a8 = (Y.Y^T) o B, a9 = (Y.Y^T) o U (NOT symmetrised),
    B / U = reciprocal / one-way part of S."""

    def __init__(self, S: CSRGraph, Y: CSRGraph):
        if S.n_rows != S.n_cols or Y.n_rows != S.n_rows:
            raise ValueError("MotifTerms: S must be [U, U] and Y [U, I]")
        n, dev = S.n_rows, S.device
        self.n, self.device, self.S, self.Y = n, dev, S, Y
        St = S.transpose()
        b_vals = S.vals * csr_sample(St, S)                       # B = S o S^T on S's pattern          (mhcn.py:343)
        self.B = B = from_coo_sum([to_coo(S, b_vals)], n, n, dev)
        self.U = U = from_coo_sum([to_coo(S, S.vals - b_vals)], n, n, dev)   # U = S - B                 (mhcn.py:344)
        Ut = U.transpose()
        Bt = B.transpose()                                        # B is symmetric; kept general

        def term(p: CSRGraph, qt: CSRGraph, m: CSRGraph) -> Coo:  # (P . Q) o M with Q given by its transpose
            return to_coo(m, masked_product(p, qt, m))

        sym = lambda t: [t, transpose_coo(t)]                     # C + C^T
        c1 = [term(U, Ut, Ut)]                                                   # (U.U) o U^T                     :345
        c2 = [term(B, Ut, Ut), term(U, Bt, Ut), term(U, Ut, B)]                  #                                 :347
        c3 = [term(B, Bt, U), term(B, Ut, B), term(U, Bt, B)]                    #                                 :349
        a4 = [term(B, Bt, B)]                                                    #                                 :351
        c5 = [term(U, Ut, U), term(U, U, U), term(Ut, Ut, U)]                    # (U.U)oU + (U.U^T)oU + (U^T.U)oU :352
        a6 = [term(U, Bt, U), term(B, U, Ut), term(Ut, Ut, B)]                   # (U.B)oU + (B.U^T)oU^T + (U^T.U)oB :354
        a7 = [term(Ut, Bt, Ut), term(B, Ut, U), term(U, U, B)]                   # (U^T.B)oU^T + (B.U)oU + (U.U^T)oB :355
        self.social: List[Coo] = a4 + a6 + a7
        for c in (c1, c2, c3, c5):                                               # A1, A2, A3, A5 = C + C^T        :346-353
            for t in c:
                self.social += sym(t)
        self.a8 = term(Y, Y, B)                                                  # (Y.Y^T) o B   (Q = Y^T, Q^T = Y) :356
        self.a9 = term(Y, Y, U)                                                  # (Y.Y^T) o U                     :357

    def co_purchase(self, max_products: int = MAX_PRODUCTS_PER_CHUNK) -> CSRGraph:
        """Y . Y^T (mhcn.py:359, esrf.py:1088)."""
        return spgemm(self.Y, self.Y.transpose(), max_products=max_products)


def build_hyper_adj_mats(S: CSRGraph, Y: CSRGraph, *, p_threshold: float = 3.0,
                         max_products: int = MAX_PRODUCTS_PER_CHUNK) -> List[CSRGraph]:
    """[H_s, H_j, H_p] of mhcn.py:340-368 from the directed social matrix S [U, U] and the interaction matrix Y [U, I]
    (canonical CSR operators on the GPU, `CSRGraph.from_scipy(...)` of `social_data.get_social_mat()` /
    `data.interaction_mat`).  Row-normalised like the reference; H_p keeps co-purchase counts > p_threshold."""
    t = MotifTerms(S, Y)
    n, dev = t.n, t.device
    H_s = from_coo_sum(t.social, n, n, dev, norm="row")                      # sum + row normalisation         :361-362
    a9 = [t.a9, transpose_coo(t.a9)]                                         # A9 = A9 + A9^T                  :358
    H_j = from_coo_sum([t.a8] + a9, n, n, dev, norm="row")                   #                                 :363-364
    yy = t.co_purchase(max_products)
    neg = lambda c: (c[0], c[1], -c[2])
    a10 = from_coo_sum([to_coo(yy, drop_zeros=False), neg(t.a8)] + [neg(c) for c in a9], n, n, dev)   # :359
    r, c, v = to_coo(a10)
    keep = v > p_threshold                                                   # H_p o (H_p > 3)                 :365-366
    H_p = from_coo_sum([(r[keep], c[keep], v[keep])], n, n, dev, norm="row")  #                                :367
    return [H_s, H_j, H_p]


def build_motif_induced_adjacency_matrix(S: CSRGraph, Y: CSRGraph, *, p_threshold: float = 5.0,
                                         max_products: int = MAX_PRODUCTS_PER_CHUNK) -> CSRGraph:
    """ESRF's single high-order adjacency (univariate/esrf.py:1067-1096): S + A1..A7 + A8 + A9 + A10 with A9 left one-sided,
    A10 = Y.Y^T without its diagonal and only where > p_threshold common purchases, every row divided by its sum."""
    t = MotifTerms(S, Y)
    n, dev = t.n, t.device
    r, c, v = to_coo(t.co_purchase(max_products))
    keep = (r != c) & (v > p_threshold)                                      # esrf.py:1089-1092
    a10 = (r[keep], c[keep], v[keep])
    return from_coo_sum([to_coo(S)] + t.social + [t.a8, t.a9, a10], n, n, dev, norm="row")   # esrf.py:1094-1096
