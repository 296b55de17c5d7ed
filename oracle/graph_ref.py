"""Oracle: adjacency construction (numpy, integer-exact).  Test infrastructure only.

Follows
  lightgcn.py:36-39          edge_index = [[u | i+U], [i+U | u]]
  selfcf.py:297-306          csr_matrix((1, (u, i+U))) ; adj = tmp + tmp.T   (duplicates summed, canonical CSR)
  selfcf.py:240-255          D^-1/2 A D^-1/2 (square) / D^-1 A (rectangular), inf -> 0
  ncl.py:76-85               raw COO of ones in insertion order, duplicates kept
  PyG gcn_norm (called by LGConv, lightgcn.py:25; torch_geometric is not in the reference tree):
      deg = scatter_add(ones, col) ; dis = deg^-1/2, inf -> 0 ; w = dis[row] * dis[col]
"""
from __future__ import annotations

import numpy as np


def bipartite_edge_index(users, items, n_users):
    users = np.asarray(users, dtype=np.int64)
    items = np.asarray(items, dtype=np.int64)
    return np.stack([np.concatenate([users, items + n_users]), np.concatenate([items + n_users, users])])


def coo_to_canonical_csr(rows, cols, vals, n_rows, n_cols):
    """Sort by (row, col), sum duplicates in their original order (stable) -> (row_ptr, col_idx, vals)."""
    rows = np.asarray(rows, dtype=np.int64)
    cols = np.asarray(cols, dtype=np.int64)
    vals = np.ones(rows.shape[0], dtype=np.float32) if vals is None else np.asarray(vals, dtype=np.float32)
    order = np.lexsort((cols, rows))  # stable: ties keep input order
    r, c, v = rows[order], cols[order], vals[order]
    if r.shape[0] == 0:
        return np.zeros(n_rows + 1, np.int32), np.zeros(0, np.int32), np.zeros(0, np.float32)
    head = np.ones(r.shape[0], dtype=bool)
    head[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
    seg = np.cumsum(head) - 1
    out_v = np.zeros(int(seg[-1]) + 1, dtype=np.float32)
    # sequential fp32 accumulation in original order (np.add.at applies updates in index order)
    np.add.at(out_v, seg, v)
    ur, uc = r[head], c[head]
    row_ptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(row_ptr, ur + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    return row_ptr.astype(np.int32), uc.astype(np.int32), out_v


def degrees(idx, n_nodes):
    return np.bincount(np.asarray(idx, dtype=np.int64), minlength=n_nodes).astype(np.int32)


def normalize_csr(row_ptr, col_idx, vals, n_rows, n_cols, mode):
    """mode 'sym' | 'row' | 'none' -> (vals_out fp32, rowsum fp32, dinv fp32)   (selfcf.py:240-255)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    vals = np.asarray(vals, dtype=np.float32)
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(row_ptr))
    rowsum = np.zeros(n_rows, dtype=np.float32)
    np.add.at(rowsum, row_of, vals)
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == "sym":
            if n_rows != n_cols:
                raise ValueError("sym needs a square matrix")
            dinv = np.power(rowsum, np.float32(-0.5)).astype(np.float32)
        elif mode == "row":
            dinv = np.power(rowsum, np.float32(-1.0)).astype(np.float32)
        elif mode == "none":
            return vals.copy(), rowsum, np.ones(n_rows, dtype=np.float32)
        else:
            raise ValueError(mode)
    dinv[np.isinf(dinv)] = 0.0
    out = (dinv[row_of] * vals).astype(np.float32)
    if mode == "sym":
        out = (out * dinv[np.asarray(col_idx, dtype=np.int64)]).astype(np.float32)
    return out, rowsum, dinv


def gcn_norm_weights(edge_index, n_nodes):
    """PyG gcn_norm(add_self_loops=False): per-edge weights in edge order + integer degrees."""
    row, col = np.asarray(edge_index[0], np.int64), np.asarray(edge_index[1], np.int64)
    deg = np.bincount(col, minlength=n_nodes).astype(np.float32)
    with np.errstate(divide="ignore"):
        dis = (np.float32(1.0) / np.sqrt(deg)).astype(np.float32)  # torch pow(-0.5) == 1/sqrt
    dis[np.isinf(dis)] = 0.0
    return (dis[row] * dis[col]).astype(np.float32), deg.astype(np.int32)


def csr_transpose(row_ptr, col_idx, vals, n_rows, n_cols):
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(row_ptr))
    order = np.argsort(np.asarray(col_idx, dtype=np.int64), kind="stable")
    t_row_ptr = np.zeros(n_cols + 1, dtype=np.int64)
    np.add.at(t_row_ptr, np.asarray(col_idx, dtype=np.int64) + 1, 1)
    return np.cumsum(t_row_ptr).astype(np.int32), row_of[order].astype(np.int32), np.asarray(vals, np.float32)[order]


def spmm_csr(row_ptr, col_idx, vals, x, dtype=np.float64):
    """Y = A @ X with accumulation in `dtype` (float64 by default: the 'true' value parity is judged against)."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n_rows = row_ptr.shape[0] - 1
    x = np.asarray(x)
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(row_ptr))
    contrib = np.asarray(vals, dtype=dtype)[:, None] * x[np.asarray(col_idx, dtype=np.int64)].astype(dtype)
    y = np.zeros((n_rows, x.shape[1]), dtype=dtype)
    np.add.at(y, row_of, contrib)
    return y


def propagate(row_ptr, col_idx, vals, x0, n_layers, mode="mean", dtype=np.float64):
    """[E0..EK], combined   (ncl.py:415-422 mean; lightgcn.py:21-27 sum)."""
    layers = [np.asarray(x0, dtype=dtype)]
    for _ in range(n_layers):
        layers.append(spmm_csr(row_ptr, col_idx, vals, layers[-1], dtype=dtype))
    total = np.sum(np.stack(layers), axis=0)
    return layers, (total / (n_layers + 1) if mode == "mean" else total)
