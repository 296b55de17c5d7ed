"""The reference's loss functions under their own names and signatures, on the libgcf kernels.

    bpr_loss(user_emb, pos_item_emb, neg_item_emb)                 ncl.py:116-120, mhcn.py:35-39
    l2_reg_loss(reg, *args)                                        ncl.py:122-123, directau.py:35, ssl4rec.py:16
    InfoNCE(view1, view2, temperature, b_cos=True)                 ncl.py:125-130, ssl4rec.py:19-23
    batch_softmax_loss(user_emb, item_emb, temperature)            ssl4rec.py:25-30
    info_nce_loss(z1, z2, temp=0.2)                                gcl.py:28-35
    NCLLosses(...).ssl_layer_loss / .ProtoNCE_loss                 ncl.py:358-375 (methods of NCLModel)
    DirectAULosses(gamma).alignment / .uniformity / .calculate_loss   directau.py:240-251 (methods of DirectAU)
"""
from __future__ import annotations

import torch

from . import functional as F_

l2_reg_loss = F_.l2_reg_loss


def bpr_loss(user_emb: torch.Tensor, pos_item_emb: torch.Tensor, neg_item_emb: torch.Tensor) -> torch.Tensor:
    """mean(-log(10e-6 + sigmoid(<u,p> - <u,n>)))  on already-gathered [B, d] rows."""
    return F_.bpr_loss_rows(user_emb, pos_item_emb, neg_item_emb, variant="log_eps_sigmoid", eps=10e-6, reduction="mean")


def InfoNCE(view1: torch.Tensor, view2: torch.Tensor, temperature: float, b_cos: bool = True) -> torch.Tensor:
    return F_.info_nce(view1, view2, temperature, b_cos)


def batch_softmax_loss(user_emb: torch.Tensor, item_emb: torch.Tensor, temperature: float) -> torch.Tensor:
    return F_.batch_softmax(user_emb, item_emb, temperature)


def info_nce_loss(z1: torch.Tensor, z2: torch.Tensor, temp: float = 0.2) -> torch.Tensor:
    return F_.info_nce_symmetric(z1, z2, temp)


class NCLLosses:
    """The two contrastive terms of NCLModel (ncl.py:358-375) with the attributes they read from `self`.
    Centroids and cluster assignments are inputs (faiss k-means in the reference, ncl.py:347-356)."""

    def __init__(self, user_num: int, item_num: int, ssl_temp: float, ssl_reg: float, alpha: float, proto_reg: float,
                 batch_size: int):
        self.user_num, self.item_num = user_num, item_num
        self.ssl_temp, self.ssl_reg, self.alpha, self.proto_reg, self.batch_size = ssl_temp, ssl_reg, alpha, proto_reg, batch_size
        self.user_centroids = self.user_2cluster = self.item_centroids = self.item_2cluster = None

    def ssl_layer_loss(self, context: torch.Tensor, initial: torch.Tensor, user, item) -> torch.Tensor:
        u = self.user_num
        dev = context.device
        user, item = F_._idx(user, dev, "user"), F_._idx(item, dev, "item")
        cu, ci = context[:u], context[u:]
        iu, ii = initial[:u], initial[u:]
        # B x U and B x I logits (4096 x 52,643 / 4096 x 91,599 at cfg 2) live only as tensor-core tiles
        loss_u = F_.ssl_layer_side(F_.gather_rows(cu, user), iu, user, self.ssl_temp)
        loss_i = F_.ssl_layer_side(F_.gather_rows(ci, item), ii, item, self.ssl_temp)
        return self.ssl_reg * (loss_u + self.alpha * loss_i)

    def ProtoNCE_loss(self, initial_emb: torch.Tensor, user_idx, item_idx) -> torch.Tensor:
        u = self.user_num
        dev = initial_emb.device
        user_idx, item_idx = F_._idx(user_idx, dev, "user_idx"), F_._idx(item_idx, dev, "item_idx")
        user_emb, item_emb = initial_emb[:u], initial_emb[u:]
        user2centroids = self.user_centroids.to(dev)[self.user_2cluster.to(dev)[user_idx]]
        item2centroids = self.item_centroids.to(dev)[self.item_2cluster.to(dev)[item_idx]]
        loss_user = InfoNCE(F_.gather_rows(user_emb, user_idx), user2centroids, self.ssl_temp) * self.batch_size
        loss_item = InfoNCE(F_.gather_rows(item_emb, item_idx), item2centroids, self.ssl_temp) * self.batch_size
        return self.proto_reg * (loss_user + loss_item)


class DirectAULosses:
    """alignment / uniformity / calculate_loss of DirectAU (directau.py:240-251)."""

    def __init__(self, gamma: float):
        self.gamma = gamma

    def alignment(self, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        return F_.directau_terms(x, y)[0]

    def uniformity(self, x: torch.Tensor, t: float = 2) -> torch.Tensor:
        return F_.directau_terms(x, x.detach(), t)[1]

    def calculate_loss(self, user_emb: torch.Tensor, item_emb: torch.Tensor) -> torch.Tensor:
        t3 = F_.directau_terms(user_emb, item_emb)   # one fused evaluation of all three terms
        return t3[0] + self.gamma * (t3[1] + t3[2]) / 2
