"""Drop-ins for univariate/sept.py: the edge-dropout augmentor and the SEPT encoder.

    GraphAugmentor.edge_dropout(sp_adj, drop_rate)      sept.py:53-62   exactly int(nnz (1 - rate)) entries kept, values 1
    SEPTEncoder(data, emb_size, n_layers, drop_rate)    sept.py:211-226,228-236 (build, encoder, the per-epoch augmented pass)
        .encoder(emb, adj) = mean over [E0, normalize(A E0), normalize(A normalize(A E0)), ...]

Every layer is one SpMM launch with the row-L2-normalise epilogue fused in (functional.spmm_normalize); the augmented operator
keeps the sparsity pattern of the full one and carries its exact transpose for the backward.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import functional as F_
from .encoders import _device, _graph_of
from .graph import CSRGraph


class GraphAugmentor:
    """edge_dropout(sp_adj, drop_rate): `row, col = sp_adj.nonzero()` enumerates every STORED entry (duplicate interactions
    included, sept.py:136-145 builds the adjacency as a COO with duplicates), exactly int(n (1 - drop_rate)) of them are
    kept uniformly at random without replacement, and a 0/1-valued matrix is rebuilt from the kept ones (duplicates sum).
    Here: one Philox key per stored entry, the kth-smallest key as threshold, and the CSR build kernels for the kept entries.
    Called once per epoch in SEPT.train, so the rebuild is off the per-batch path."""

    _calls = 0

    @staticmethod
    def edge_dropout(sp_adj, drop_rate: float, *, seed: int = 0) -> CSRGraph:
        dev = _device()
        lib = _lib.load()
        coo = sp_adj.tocoo()
        rows = torch.from_numpy(np.ascontiguousarray(coo.row, dtype=np.int64)).to(dev)
        cols = torch.from_numpy(np.ascontiguousarray(coo.col, dtype=np.int64)).to(dev)
        n = rows.numel()
        keep = int(n * (1 - drop_rate))
        GraphAugmentor._calls += 1
        keys = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
        _lib.check(lib.gcf_philox_keys(n, int(seed) & (2**64 - 1), GraphAugmentor._calls, _lib.ptr(keys), _lib.current_stream()),
                   "gcf_philox_keys")
        order_key = keys[:n] * max(n, 1) + torch.arange(n, device=dev)          # ties (2^-32 per pair) broken by entry id
        if keep <= 0:
            mask = torch.zeros(n, dtype=torch.bool, device=dev)
        elif keep >= n:
            mask = torch.ones(n, dtype=torch.bool, device=dev)
        else:
            mask = order_key <= torch.kthvalue(order_key, keep).values
        return CSRGraph.from_coo(rows[mask], cols[mask], None, coo.shape[0], coo.shape[1], norm="none")


class SEPTEncoder(nn.Module):
    """user_embeddings / item_embeddings: nn.Embedding (state_dict keys user_embeddings.weight, item_embeddings.weight)."""

    def __init__(self, data, emb_size: int, n_layers: int, drop_rate: float):
        super().__init__()
        dev = _device()
        self.data, self.n_layers, self.drop_rate = data, n_layers, drop_rate
        self.user_embeddings = nn.Embedding(data.user_num, emb_size).to(dev)
        self.item_embeddings = nn.Embedding(data.item_num, emb_size).to(dev)
        self.norm_adj = _graph_of(data)

    def encoder(self, emb: torch.Tensor, adj: CSRGraph) -> torch.Tensor:
        all_embs = [emb]
        for _ in range(self.n_layers):
            emb = F_.spmm_normalize(adj, emb)      # torch.sparse.mm + F.normalize(dim=1), one launch
            all_embs.append(emb)
        return torch.stack(all_embs, dim=0).mean(0)

    def forward(self):
        """The per-epoch pass of SEPT.train (sept.py:232-236): augmented graph -> (U, V)."""
        all_emb = torch.cat([self.user_embeddings.weight, self.item_embeddings.weight], dim=0)
        adj = GraphAugmentor.edge_dropout(self.data.norm_adj, self.drop_rate) if self.training else self.norm_adj
        final = self.encoder(all_emb, adj)
        return final[: self.data.user_num], final[self.data.user_num:]
