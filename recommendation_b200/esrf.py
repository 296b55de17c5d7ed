"""Drop-in for the model part of univariate/esrf.py (ESRF: adversarial social recommendation).

    build_motif_induced_adjacency_matrix(S, Y)   esrf.py:1067-1096  (motifs.py: masked sparse products + expand-sort-compress)
    create_joint_sparse_adjacency(users, items)  esrf.py:955-971    D^-1/2 (R + R^T) D^-1/2 on the GPU build kernels
    gumbel_softmax(logits, temperature, u=None)  esrf.py:1003-1008  (u: optional uniform noise, injected by parity runs)
    Generator, Discriminator                     esrf.py:1116-1198  same parameters / state_dict keys / forward signatures
    pairwise_losses / adversarial_losses         the loss lines of trainModel (esrf.py:1231-1236, 1296-1309)

Sparse propagation layers are one SpMM launch with the row-L2-normalise epilogue fused in (the raw product feeds the next
layer, the normalised one the layer combination -- both written by the same launch); the BPR-style losses use the fused
gather + loss kernel (sum reduction, eps 1e-10).  The alternative neighbourhood is a dense [U, U] matrix in the reference and
stays a dense cuBLAS product here (`torch.mm(alternative_neighborhood, ...)`, esrf.py:1176, 1300).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F_
from . import motifs
from .encoders import _device
from .graph import CSRGraph

build_motif_induced_adjacency_matrix = motifs.build_motif_induced_adjacency_matrix


def create_joint_sparse_adjacency(users: torch.Tensor, items: torch.Tensor, num_users: int, num_items: int) -> CSRGraph:
    """esrf.py:955-971: csr((1, (u, i + U))) with duplicate ratings summed, A = R + R^T, D^-1/2 A D^-1/2."""
    return CSRGraph.from_pairs(users, items, num_users, num_items, norm="sym")


def gumbel_softmax(logits: torch.Tensor, temperature: float = 0.2, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    eps = 1e-10
    if u is None:
        u = torch.rand_like(logits)
    gumbel_noise = -torch.log(-torch.log(u + eps) + eps)
    return TF.softmax((torch.log(logits + eps) + gumbel_noise) / temperature, dim=-1)


class Generator(nn.Module):
    """esrf.py:1116-1149.  state_dict keys: relation_embeddings, projection_head, c_selector."""

    def __init__(self, num_users: int, emb_size: int, n_layers: int, K: int):
        super().__init__()
        dev = _device()
        self.relation_embeddings = nn.Parameter(torch.randn(num_users, emb_size, device=dev) * 0.005)
        self.projection_head = nn.Parameter(torch.randn(emb_size, emb_size, device=dev) * 0.005)
        self.c_selector = nn.Parameter(torch.randn(K, num_users, device=dev) * 0.005)
        self.n_layers, self.K, self.num_users, self.emb_size = n_layers, K, num_users, emb_size

    def propagate(self, A: CSRGraph) -> torch.Tensor:
        """mean over [E0, normalize(A E0), normalize(A A E0), ...]: the RAW product feeds the next layer (esrf.py:1128-1134)."""
        all_embeddings = [self.relation_embeddings]
        user_embeddings = self.relation_embeddings
        for _ in range(self.n_layers):
            user_embeddings, norm_embeddings = F_.spmm_and_normalize(A, user_embeddings)
            all_embeddings.append(norm_embeddings)
        return torch.stack(all_embeddings, dim=0).mean(dim=0)

    def forward(self, A: CSRGraph, user_segment: int, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """noise: optional uniform draws [segment, K, num_users] replacing torch.rand_like (parity runs)."""
        user_embeddings = self.propagate(A)
        segment_end = min(user_segment + 100, self.num_users)
        user_features = torch.mm(user_embeddings[user_segment:segment_end], user_embeddings.t())
        alpha = user_features.unsqueeze(1) * self.c_selector.unsqueeze(0)          # [segment, K, U]: all rows of the loop at once
        segment = gumbel_softmax(alpha, 0.2, noise).sum(dim=1)
        alternative_neighborhood = torch.zeros(self.num_users, self.num_users, device=user_embeddings.device)
        alternative_neighborhood[user_segment:segment_end] = segment
        return alternative_neighborhood


class Discriminator(nn.Module):
    """esrf.py:1151-1198.  state_dict keys: user_embeddings, item_embeddings,
    attention_weights.{k}.attention_m1{k}.weight / attention_m2{k}.weight / attention_v{k}.weight (unused by forward there too)."""

    def __init__(self, num_users: int, num_items: int, emb_size: int, n_layers: int):
        super().__init__()
        dev = _device()
        self.user_embeddings = nn.Parameter(torch.randn(num_users, emb_size, device=dev) * 0.01)
        self.item_embeddings = nn.Parameter(torch.randn(num_items, emb_size, device=dev) * 0.01)
        self.attention_weights = nn.ModuleList()
        for k in range(n_layers):
            self.attention_weights.append(nn.ModuleDict({
                f"attention_m1{k}": nn.Linear(emb_size, emb_size, bias=False),
                f"attention_m2{k}": nn.Linear(emb_size, emb_size, bias=False),
                f"attention_v{k}": nn.Linear(emb_size * 2, 1, bias=False),
            }).to(dev))
        self.n_layers, self.num_users, self.num_items, self.emb_size = n_layers, num_users, num_items, emb_size

    def forward(self, norm_adj: CSRGraph, alternative_neighborhood: Optional[torch.Tensor], is_social, is_attentive, K
                ) -> Tuple[torch.Tensor, torch.Tensor]:
        ego = torch.cat([self.user_embeddings, self.item_embeddings], dim=0)
        all_embeddings = [ego]
        for _ in range(self.n_layers):
            if is_social:
                # the reference computes the sparse product in this branch too and discards it (esrf.py:1173-1186)
                social = torch.mm(alternative_neighborhood, ego[: self.num_users]) / K
                ego = torch.cat([ego[: self.num_users] + social, ego[self.num_users:]], dim=0)
                norm_embeddings = TF.normalize(ego, p=2, dim=1)
            else:
                ego, norm_embeddings = F_.spmm_and_normalize(norm_adj, ego)      # ego = A ego, normalize(A ego): one launch
            all_embeddings.append(norm_embeddings)
        total = torch.stack(all_embeddings, dim=0).sum(dim=0)
        return torch.split(total, [self.num_users, self.num_items])


def pairwise_losses(user_emb: torch.Tensor, item_emb: torch.Tensor, user_idx, i_idx, j_idx, reg_u: float):
    """(pairwise_loss, reg_loss) of esrf.py:1231-1236: -sum log(sigmoid(y_ui - y_uj) + 1e-10) and
    regU (|u_emb|_F + |v_emb|_F + |neg_emb|_F) over the gathered rows (Frobenius norms, not squared)."""
    pairwise = F_.bpr_loss_gather(user_emb, item_emb, user_idx, i_idx, j_idx, variant="log_eps_sigmoid", eps=1e-10, reduction="sum")
    u, v, n = F_.gather_rows(user_emb, user_idx), F_.gather_rows(item_emb, i_idx), F_.gather_rows(item_emb, j_idx)
    return pairwise, reg_u * (torch.norm(u) + torch.norm(v) + torch.norm(n))


def adversarial_losses(user_emb: torch.Tensor, item_emb: torch.Tensor, alternative_neighborhood: torch.Tensor, user_idx, i_idx,
                       K: int):
    """(adversarial_loss of the discriminator, g_adv_loss of the generator), esrf.py:1296-1309."""
    dev = user_emb.device
    user_idx = torch.as_tensor(user_idx, dtype=torch.int64, device=dev)
    u_emb, v_emb = F_.gather_rows(user_emb, user_idx), F_.gather_rows(item_emb, i_idx)
    y_ui = torch.sum(u_emb * v_emb, dim=1)
    friend = torch.mm(alternative_neighborhood[user_idx], user_emb) / K
    y_vi = torch.sum(friend * v_emb, dim=1)
    d_adv = -torch.sum(torch.log(torch.sigmoid(y_ui - y_vi) + 1e-10))
    g_adv = -torch.sum(torch.log(torch.sigmoid(y_vi - y_ui) + 1e-10))
    return d_adv, g_adv
