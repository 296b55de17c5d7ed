"""The NCL training iteration (ncl.py:308-329) assembled from the drop-in pieces, plus the E-step (ncl.py:339-356).

    model = LGCNEncoder(data, emb_size, n_layers); ncl = NCLLosses(...)
    e_step(model, ncl, k)                                   # k-means of the propagated embeddings (faiss in the reference)
    for batch in next_batch_pairwise(data, batch_size):
        total, parts = ncl_step(model, ncl, optimizer, batch, reg, batch_size, hyper_layers)
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import functional as F_
from .encoders import LGCNEncoder
from .kmeans import run_kmeans
from .losses import NCLLosses, bpr_loss, l2_reg_loss


@torch.no_grad()
def e_step(model: LGCNEncoder, ncl: NCLLosses, k: int, *, seed: int = 1234) -> int:
    """ncl.py:339-345: cluster the propagated user and item embeddings; centroids / assignments land on `ncl`."""
    user_emb, item_emb, _ = model()
    ncl.user_centroids, ncl.user_2cluster, k = run_kmeans(user_emb, k, seed=seed)
    ncl.item_centroids, ncl.item_2cluster, k = run_kmeans(item_emb, k, seed=seed + 1)
    return k


def ncl_step(model: LGCNEncoder, ncl: NCLLosses, optimizer, batch, reg: float, batch_size: int, hyper_layers: int,
             *, k: int = 0, refresh_clusters: bool = False) -> Tuple[torch.Tensor, Dict[str, torch.Tensor]]:
    """One iteration of the batch loop (ncl.py:313-329).  `refresh_clusters` re-runs the E-step between the structural and
    the prototype loss exactly where the reference does (ncl.py:324)."""
    user_idx, pos_idx, neg_idx = batch
    model.train()
    rec_user_emb, rec_item_emb, emb_list = model()
    user_emb = F_.gather_rows(rec_user_emb, user_idx)
    pos_emb = F_.gather_rows(rec_item_emb, pos_idx)
    neg_emb = F_.gather_rows(rec_item_emb, neg_idx)
    rec_loss = bpr_loss(user_emb, pos_emb, neg_emb)
    initial_emb = emb_list[0]
    context_emb = emb_list[-1] if hyper_layers * 2 >= len(emb_list) else emb_list[hyper_layers * 2]
    ssl_loss = ncl.ssl_layer_loss(context_emb, initial_emb, user_idx, pos_idx)
    if refresh_clusters:
        e_step(model, ncl, k)
    proto_loss = ncl.ProtoNCE_loss(initial_emb, user_idx, pos_idx)
    total = rec_loss + l2_reg_loss(reg, user_emb, pos_emb, neg_emb) / batch_size + ssl_loss + proto_loss
    optimizer.zero_grad()
    total.backward()
    optimizer.step()
    return total, {"rec": rec_loss, "ssl": ssl_loss, "proto": proto_loss}
