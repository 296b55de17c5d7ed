"""Drop-ins for univariate/buir.py: the LGCN encoder with stochastic edge dropout and the BUIR_NB model.

    LGCN_Encoder(data, emb_size, n_layers, drop_rate, drop_flag=False)   buir.py:280-341
        .forward(inputs) -> (user_all[users], item_all[items]);  .get_embedding() -> (user_all, item_all)
    BUIR_NB(data, emb_size, momentum, n_layers, drop_rate, drop_flag=False)   buir.py:236-277

sparse_dropout (buir.py:300-309) keeps every stored adjacency entry with probability 1 - rate, rescales by 1 / (1 - rate)
and uses the same dropped operator for all layers of one forward; here the mask comes from the Philox counter of the entry
(`CSRGraph.dropout`, gcf_csr_dropout_values), the dropped operator carries its exact transpose for the backward, and the
K layers + mean run in the fused SpMM kernels.  The rate itself is `np.random.random() * drop_ratio` as in the reference.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F_
from .encoders import _LGCNBase, _device


class LGCN_Encoder(_LGCNBase):
    def __init__(self, data, emb_size: int, n_layers: int, drop_rate: float, drop_flag: bool = False, *, seed: int = 0):
        super().__init__(data, emb_size, n_layers)
        self.drop_ratio = drop_rate
        self.drop_flag = drop_flag
        self._seed, self._calls = seed, 0

    def _propagate(self, graph):
        final = F_.propagate(graph, self.joint_table(), self.layers, mode="mean")
        u = self.data.user_num
        return final[:u], final[u:]

    def forward(self, inputs):
        graph = self.sparse_norm_adj
        if self.drop_flag:
            self._calls += 1
            graph = graph.dropout(np.random.random() * self.drop_ratio, seed=self._seed, offset=self._calls)
        user_all, item_all = self._propagate(graph)
        dev = user_all.device
        users, items = F_._idx(inputs["user"], dev, "user"), F_._idx(inputs["item"], dev, "item")
        return F_.gather_rows(user_all, users), F_.gather_rows(item_all, items)

    @torch.no_grad()
    def get_embedding(self):
        return self._propagate(self.sparse_norm_adj)


class BUIR_NB(nn.Module):
    def __init__(self, data, emb_size: int, momentum: float, n_layers: int, drop_rate: float, drop_flag: bool = False):
        super().__init__()
        self.emb_size = emb_size
        self.momentum = momentum
        self.online_encoder = LGCN_Encoder(data, emb_size, n_layers, drop_rate, drop_flag, seed=1)
        self.target_encoder = LGCN_Encoder(data, emb_size, n_layers, drop_rate, drop_flag, seed=2)
        self.predictor = nn.Linear(emb_size, emb_size).to(_device())
        self._init_target()

    def _init_target(self):
        for param_o, param_t in zip(self.online_encoder.parameters(), self.target_encoder.parameters()):
            param_t.data.copy_(param_o.data)
            param_t.requires_grad = False

    def update_target(self, u_idx, i_idx):
        on, tg = self.online_encoder.embedding_dict, self.target_encoder.embedding_dict
        dev = on["user_emb"].device
        u_idx, i_idx = F_._idx(u_idx, dev, "u_idx"), F_._idx(i_idx, dev, "i_idx")
        tg["user_emb"].data[u_idx] = tg["user_emb"].data[u_idx] * self.momentum + on["user_emb"].data[u_idx] * (1 - self.momentum)
        tg["item_emb"].data[i_idx] = tg["item_emb"].data[i_idx] * self.momentum + on["item_emb"].data[i_idx] * (1 - self.momentum)

    def forward(self, inputs):
        u_online, i_online = self.online_encoder(inputs)
        u_target, i_target = self.target_encoder(inputs)
        return self.predictor(u_online), u_target, self.predictor(i_online), i_target

    @torch.no_grad()
    def get_embedding(self):
        u_online, i_online = self.online_encoder.get_embedding()
        return self.predictor(u_online), u_online, self.predictor(i_online), i_online

    def get_loss(self, output):
        u_online, u_target, i_online, i_target = output
        u_online, u_target = TF.normalize(u_online, dim=-1), TF.normalize(u_target, dim=-1)
        i_online, i_target = TF.normalize(i_online, dim=-1), TF.normalize(i_target, dim=-1)
        loss_ui = 2 - 2 * (u_online * i_target).sum(dim=-1)
        loss_iu = 2 - 2 * (i_online * u_target).sum(dim=-1)
        return (loss_ui + loss_iu).mean()
