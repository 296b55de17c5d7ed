"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement (scipy, fp32 like the reference) of the motif-induced adjacency matrices of MHCN,
univariate/mhcn.py:340-368 (`build_hyper_adj_mats`).  Pinned by tests/golden/mhcn_motifs.npz, which holds the outputs of
the reference's own function on the same S / Y (tests/test_oracle_golden.py::test_motif_oracle_matches_reference_fixture).

Notation of the reference: S directed social matrix (S[a, b] = 1: a follows b), Y user-item matrix, B = S o S^T the
reciprocal part, U = S - B the one-way part ("o" = element-wise product, "." = matrix product).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp


def masked_product(P, Q, M):
    """(P . Q) o M -- the shape every motif term has (mhcn.py:345-360)."""
    return (P @ Q).multiply(M).tocsr()


def motif_terms(S, Y):
    """A1..A10 of mhcn.py:345-361 as a dict (before the channel sums)."""
    S = sp.csr_matrix(S, dtype=np.float32)
    Y = sp.csr_matrix(Y, dtype=np.float32)
    B = S.multiply(S.T).tocsr()                    # mhcn.py:343
    U = (S - B).tocsr()                            # mhcn.py:344
    U.eliminate_zeros()
    Ut = U.T.tocsr()
    mp = masked_product
    c1 = mp(U, U, Ut)                                              # :345
    c2 = mp(B, U, Ut) + mp(U, B, Ut) + mp(U, U, B)                 # :347
    c3 = mp(B, B, U) + mp(B, U, B) + mp(U, B, B)                   # :349
    c5 = mp(U, U, U) + mp(U, Ut, U) + mp(Ut, U, U)                 # :352
    a = {
        "A1": c1 + c1.T, "A2": c2 + c2.T, "A3": c3 + c3.T,         # :346,348,350
        "A4": mp(B, B, B),                                         # :351
        "A5": c5 + c5.T,                                           # :353
        "A6": mp(U, B, U) + mp(B, Ut, Ut) + mp(Ut, U, B),          # :354
        "A7": mp(Ut, B, Ut) + mp(B, U, U) + mp(U, Ut, B),          # :355
    }
    yy = (Y @ Y.T).tocsr()
    a["A8"] = yy.multiply(B).tocsr()                               # :356
    a9 = yy.multiply(U).tocsr()                                    # :357
    a["A9"] = (a9 + a9.T).tocsr()                                  # :358
    a["A10"] = (yy - a["A8"] - a["A9"]).tocsr()                    # :359
    return {k: sp.csr_matrix(v, dtype=np.float32) for k, v in a.items()}, B, U


def row_normalise(H):
    """H o (1 / rowsum) (mhcn.py:362,364,367): rows without entries stay empty."""
    H = sp.csr_matrix(H, dtype=np.float32)
    rs = np.asarray(H.sum(axis=1), dtype=np.float32).ravel()
    with np.errstate(divide="ignore"):
        inv = (np.float32(1.0) / rs).astype(np.float32)
    inv[~np.isfinite(inv)] = 0.0
    out = sp.diags(inv).dot(H).tocsr().astype(np.float32)
    out.eliminate_zeros()
    return out


def build_hyper_adj_mats(S, Y, p_threshold: float = 3.0):
    """[H_s, H_j, H_p] (mhcn.py:361-368): social / joint / purchase channels, row-normalised; H_p keeps co-purchase
    counts > p_threshold only (`H_p.multiply(H_p > 3)`)."""
    a, _, _ = motif_terms(S, Y)
    hs = sum(a[k] for k in ("A1", "A2", "A3", "A4", "A5", "A6", "A7"))
    hj = a["A8"] + a["A9"]
    hp = a["A10"].multiply(a["A10"] > p_threshold)
    return [row_normalise(hs), row_normalise(hj), row_normalise(hp)]


def build_motif_induced_adjacency_matrix(S, Y, p_threshold: float = 5.0):
    """ESRF's single high-order adjacency, univariate/esrf.py:1067-1096: S + A1..A7 + A8 + A9 (one-sided) + A10 with
    A10 = Y.Y^T minus its diagonal, kept where > p_threshold; rows divided by their sums."""
    a, B, U = motif_terms(S, Y)
    S = sp.csr_matrix(S, dtype=np.float32)
    Y = sp.csr_matrix(Y, dtype=np.float32)
    yy = (Y @ Y.T).tocsr()
    a9 = yy.multiply(U).tocsr()                              # esrf.py:1087 (not symmetrised there)
    a10 = yy - sp.diags(yy.diagonal())                       # esrf.py:1089-1090
    a10 = a10.multiply(a10 > p_threshold)                    # esrf.py:1092
    total = S + sum(a[k] for k in ("A1", "A2", "A3", "A4", "A5", "A6", "A7", "A8")) + a9 + a10
    return row_normalise(total)                              # esrf.py:1095-1096
