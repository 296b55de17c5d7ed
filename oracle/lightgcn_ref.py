"""Oracle: the lightgcn.py model and training step on torch-CPU.  Test infrastructure only.

lightgcn.py imports torch_geometric.nn.LGConv, which is not installed and not vendored in the reference tree
(version unpinned; nearest hint is the commented `pip install torch-geometric` next to torch==2.2.0+cu118 in
univariate/grace.py:1-8).  LGConv is restated here from PyG's published semantics:

    LGConv.forward(x, edge_index):  edge_index, w = gcn_norm(edge_index, None, N, improved=False, add_self_loops=False)
                                    out[col] += w_e * x[row]     (aggr='add', flow source_to_target)
    gcn_norm:  deg = scatter_add(ones(E), col, N); dis = deg.pow(-0.5); dis[dis == inf] = 0; w = dis[row] * dis[col]

and the model / step follow lightgcn.py:21-27 (x = sum_k E(k)) and lightgcn.py:83-120.  Cross-check: on the
bidirectional edge lists lightgcn.py builds, this equals selfcf.LGCN_Encoder's sym-normalised torch.sparse.mm
(tests/test_oracle_golden.py::test_lgconv_matches_selfcf_fixture), which IS importable and pinned by fixtures.
This restatement is also the timed "reference" arm of bench.py (`--impl reference`, cpu_baseline kind "port").
"""
from __future__ import annotations

import torch


def lgconv(x: torch.Tensor, edge_index: torch.Tensor) -> torch.Tensor:
    """One LGConv layer, recomputing gcn_norm on every call exactly like the reference (lightgcn.py:25)."""
    n = x.shape[0]
    row, col = edge_index[0], edge_index[1]
    deg = torch.zeros(n, dtype=x.dtype).scatter_add_(0, col, torch.ones(col.shape[0], dtype=x.dtype))
    dis = deg.pow(-0.5)
    dis = dis.masked_fill(dis == float("inf"), 0.0)
    w = dis[row] * dis[col]
    msg = w.unsqueeze(1) * x.index_select(0, row)
    return torch.zeros_like(x).index_add_(0, col, msg)


def lightgcn_forward(user_w: torch.Tensor, item_w: torch.Tensor, edge_index: torch.Tensor, n_layers: int):
    """lightgcn.py:21-27: x = cat(W_u, W_i); out = x; for conv: out = conv(out); x += out."""
    x = torch.cat([user_w, item_w], dim=0)
    out = x
    for _ in range(n_layers):
        out = lgconv(out, edge_index)
        x = x + out
    return x[: user_w.shape[0]], x[user_w.shape[0]:]


def lightgcn_step_loss(user_w, item_w, edge_index, pos_u, pos_i, neg_i, n_layers: int, reg_weight: float):
    """Forward + BPR + reg of one epoch iteration (lightgcn.py:85-118)."""
    from .losses_ref import bpr_lightgcn

    user_emb, item_emb = lightgcn_forward(user_w, item_w, edge_index, n_layers)
    return bpr_lightgcn(user_emb, item_emb, pos_u, pos_i, neg_i, reg_weight)


class SparseLightGCN(torch.nn.Module):
    """The same model with the propagation done the way the *other* reference encoders do it
    (selfcf.py:475-485: torch.sparse.mm on the COO tensor built by convert_sparse_mat_to_tensor, selfcf.py:219-225).
    This is the CPU path SURVEY.md 6 / BASELINE.md 4 name as the official CPU baseline."""

    def __init__(self, n_users: int, n_items: int, d: int, n_layers: int, norm_adj_scipy):
        super().__init__()
        self.n_users, self.n_layers = n_users, n_layers
        self.user_embedding = torch.nn.Embedding(n_users, d)
        self.item_embedding = torch.nn.Embedding(n_items, d)
        torch.nn.init.xavier_uniform_(self.user_embedding.weight)
        torch.nn.init.xavier_uniform_(self.item_embedding.weight)
        coo = norm_adj_scipy.tocoo()
        idx = torch.stack([torch.from_numpy(coo.row).long(), torch.from_numpy(coo.col).long()])
        self.adj = torch.sparse_coo_tensor(idx, torch.from_numpy(coo.data).float(), coo.shape)

    def forward(self):
        x = torch.cat([self.user_embedding.weight, self.item_embedding.weight], 0)
        out = x
        for _ in range(self.n_layers):
            out = torch.sparse.mm(self.adj, out)
            x = x + out
        return x[: self.n_users], x[self.n_users:]
