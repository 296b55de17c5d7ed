"""Drop-ins for the social-graph variants (BASELINE cfg 4): MHCN and DiffNet on the libgcf SpMM.

    MHCNModel  -- the nn.Module half of univariate/mhcn.py's MHCN: parameters of build() (mhcn.py:372-402), self_gating,
                  self_supervised_gating, channel_attention, forward(u_idx, v_idx, neg_idx) -> 6-tuple (mhcn.py:404-478),
                  hierarchical_self_supervision (mhcn.py:480-505).  The motif matrices H_s / H_j / H_p (built by sparse
                  matrix products in build_hyper_adj_mats, mhcn.py:340-368 -- out of scope) and the row-normalised R are
                  INPUTS (scipy sparse).
    DiffNetModel -- forward() of univariate/diffnet.py:1124-1132: K x (S U -> concat -> GEMM -> ReLU) + A V.

Every torch.sparse.mm of the reference is a CSR SpMM launch here (non-symmetric operators: the backward runs on the
cached transposed CSR); the dense gating / attention GEMMs stay on cuBLAS exactly as in the reference.
state_dict keys equal the reference's.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F_
from .encoders import _device
from .graph import CSRGraph


def _as_graph(mat, dev) -> CSRGraph:
    return mat if isinstance(mat, CSRGraph) else CSRGraph.from_scipy(mat, norm="none", device=dev)


class MHCNModel(nn.Module):
    def __init__(self, user_num: int, item_num: int, emb_size: int, n_layers: int, ss_rate: float, H_s, H_j, H_p, R):
        super().__init__()
        dev = _device()
        self.n_layers, self.ss_rate, self.emb_size, self.n_channel = n_layers, ss_rate, emb_size, 4
        self.user_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(user_num, emb_size, device=dev)))
        self.item_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(item_num, emb_size, device=dev)))
        self.gating_weights, self.gating_bias = nn.ParameterDict(), nn.ParameterDict()
        self.sgating_weights, self.sgating_bias = nn.ParameterDict(), nn.ParameterDict()
        for i in range(self.n_channel):
            c = str(i + 1)
            self.gating_weights[c] = nn.Parameter(nn.init.xavier_uniform_(torch.empty(emb_size, emb_size, device=dev)))
            self.gating_bias[c] = nn.Parameter(torch.zeros(1, emb_size, device=dev))
            self.sgating_weights[c] = nn.Parameter(nn.init.xavier_uniform_(torch.empty(emb_size, emb_size, device=dev)))
            self.sgating_bias[c] = nn.Parameter(torch.zeros(1, emb_size, device=dev))
        self.attention = nn.Parameter(nn.init.xavier_uniform_(torch.empty(1, emb_size, device=dev)))
        self.attention_mat = nn.Parameter(nn.init.xavier_uniform_(torch.empty(emb_size, emb_size, device=dev)))
        self.H_s, self.H_j, self.H_p, self.R = (_as_graph(m, dev) for m in (H_s, H_j, H_p, R))
        self.R_t = self.R.transpose()   # R^T as its own CSR: torch.sparse.mm(self.R.transpose(0, 1), .) of mhcn.py:452

    def self_gating(self, em, channel):
        c = str(channel)
        return torch.multiply(em, torch.sigmoid(torch.matmul(em, self.gating_weights[c]) + self.gating_bias[c]))

    def self_supervised_gating(self, em, channel):
        c = str(channel)
        return torch.multiply(em, torch.sigmoid(torch.matmul(em, self.sgating_weights[c]) + self.sgating_bias[c]))

    def channel_attention(self, *channel_embeddings):
        weights = [torch.sum(torch.multiply(self.attention, torch.matmul(e, self.attention_mat)), 1) for e in channel_embeddings]
        score = TF.softmax(torch.stack(weights), dim=0)
        mixed = 0
        for i in range(len(weights)):
            mixed = mixed + torch.mul(score[i].view(-1, 1), channel_embeddings[i])
        return mixed, score

    def forward(self, u_idx, v_idx, neg_idx, perms: Optional[Sequence[torch.Tensor]] = None):
        """perms: optional 9 row permutations (3 per hierarchical_self_supervision call, in call order) replacing
        the torch.randperm draws -- parity runs inject the reference's draws; training leaves it None."""
        c1, c2, c3 = (self.self_gating(self.user_embeddings, k) for k in (1, 2, 3))
        simple = self.self_gating(self.user_embeddings, 4)
        all_c1, all_c2, all_c3, all_simple = [c1], [c2], [c3], [simple]
        item = self.item_embeddings
        all_i = [item]
        for _ in range(self.n_layers):
            mixed, _ = self.channel_attention(c1, c2, c3)
            mixed = mixed + simple / 2
            # every torch.sparse.mm + F.normalize pair of mhcn.py:440-457 is one launch (raw product + normalised rows)
            c1, n1 = F_.spmm_and_normalize(self.H_s, c1); all_c1.append(n1)
            c2, n2 = F_.spmm_and_normalize(self.H_j, c2); all_c2.append(n2)
            c3, n3 = F_.spmm_and_normalize(self.H_p, c3); all_c3.append(n3)
            new_item, ni = F_.spmm_and_normalize(self.R_t, mixed); all_i.append(ni)
            simple, ns = F_.spmm_and_normalize(self.R, item); all_simple.append(ns)
            item = new_item
        c1, c2, c3 = (torch.stack(a).sum(dim=0) for a in (all_c1, all_c2, all_c3))
        simple = torch.stack(all_simple).sum(dim=0)
        final_item = torch.stack(all_i).sum(dim=0)
        final_user, _ = self.channel_attention(c1, c2, c3)
        final_user = final_user + simple / 2
        ss_loss = 0
        for k, adj in ((1, self.H_s), (2, self.H_j), (3, self.H_p)):
            p = None if perms is None else perms[3 * (k - 1):3 * k]
            ss_loss = ss_loss + self.hierarchical_self_supervision(self.self_supervised_gating(final_user, k), adj, p)
        ss_loss = self.ss_rate * ss_loss
        dev = final_user.device
        u_idx, v_idx, neg_idx = (F_._idx(t, dev, "idx") for t in (u_idx, v_idx, neg_idx))
        return (F_.gather_rows(final_user, u_idx), F_.gather_rows(final_item, v_idx), F_.gather_rows(final_item, neg_idx),
                ss_loss, final_user, final_item)

    def hierarchical_self_supervision(self, em, adj, perms: Optional[Sequence[torch.Tensor]] = None):
        n = em.size(0)
        p = list(perms) if perms is not None else [torch.randperm(n, device=em.device) for _ in range(3)]
        score = lambda a, b: torch.sum(torch.multiply(a, b), 1)
        edge = F_.spmm(_as_graph(adj, em.device), em)
        pos = score(em, edge)
        neg1 = score(em[p[0]], edge)
        neg2 = score(edge[p[1]], em)
        local_loss = torch.sum(-torch.log(torch.sigmoid(pos - neg1)) - torch.log(torch.sigmoid(neg1 - neg2)))
        graph = torch.mean(edge, 0, keepdim=True)
        pos = score(edge, graph.expand_as(edge))
        neg1 = score(edge[p[2]], graph.expand_as(edge))
        global_loss = torch.sum(-torch.log(torch.sigmoid(pos - neg1)))
        return global_loss + local_loss


class DiffNetModel(nn.Module):
    """user_embeddings / item_embeddings ~ randn * 0.005 (diffnet.py:1066-1067), weights: K x [2d, d] xavier (diffnet.py:1086-1089);
    S = row-normalised social matrix, A = user x item rating matrix, both scipy sparse inputs."""

    def __init__(self, num_users: int, num_items: int, emb_size: int, n_layers: int, S, A):
        super().__init__()
        dev = _device()
        self.n_layers, self.emb_size = n_layers, emb_size
        self.user_embeddings = nn.Parameter(torch.randn(num_users, emb_size, device=dev) * 0.005)
        self.item_embeddings = nn.Parameter(torch.randn(num_items, emb_size, device=dev) * 0.005)
        self.weights = nn.ParameterList([nn.Parameter(nn.init.xavier_uniform_(torch.empty(2 * emb_size, emb_size, device=dev)))
                                         for _ in range(n_layers)])
        self.S, self.A = _as_graph(S, dev), _as_graph(A, dev)

    def forward(self) -> torch.Tensor:
        user = self.user_embeddings
        for k in range(self.n_layers):
            new_user = F_.spmm(self.S, user)
            user = torch.relu(torch.matmul(torch.cat([new_user, user], dim=1), self.weights[k]))
        return user + F_.spmm(self.A, self.item_embeddings)

    def bpr_sum_loss(self, final_user: torch.Tensor, user_idx, i_idx, j_idx, regU: float) -> torch.Tensor:
        """The loss of diffnet.py:1107-1115: -sum log sigmoid(y) + regU * (|u| + |v| + |n|) (norms not squared)."""
        dev = final_user.device
        user_idx, i_idx, j_idx = (F_._idx(t, dev, "idx") for t in (user_idx, i_idx, j_idx))
        u = F_.gather_rows(final_user, user_idx)
        v = F_.gather_rows(self.item_embeddings, i_idx)
        n = F_.gather_rows(self.item_embeddings, j_idx)
        rec = F_.bpr_loss_rows(u, v, n, variant="softplus", reduction="sum")
        return rec + regU * (torch.norm(u, 2) + torch.norm(v, 2) + torch.norm(n, 2))
