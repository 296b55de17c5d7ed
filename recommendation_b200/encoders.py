"""Drop-ins for the reference's LightGCN-style encoders and the models built on them.

    LGCNEncoder(data, emb_size, n_layers).forward() -> (user[U,d], item[I,d], all_emb: list of K+1 [N,d])
                                                        ncl.py:397-422 == directau.py:269-293
    LGCN_Encoder(data, emb_size, n_layers).forward() -> (user_all, item_all)          selfcf.py:457-485
    SelfCF_HE(data, emb_size, momentum, n_layers)                                      selfcf.py:488-525
    DNNEncoder(data, emb_size, drop_rate, tau, n_layers)                               ssl4rec.py:162-196
    GRACEModel(num_users, num_items, emb_size=64, num_layers=2, proj_dim=64), EdgeRemoving(pe)   gcl.py:18-64

`data` is the reference's `Interaction` object (or anything with .user_num, .item_num, .norm_adj as a scipy sparse
matrix).  state_dict keys equal the reference's, so parameters copy over one to one.
The adjacency is canonicalised once into a device CSR (integer kernels); the propagation, the layer mean and their
backward run in the fused SpMM kernels; gathers use the warp-aggregated scatter-add backward.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import functional as F_
from .graph import CSRGraph
from .tables import JointEmbeddingDict


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("recommendation_b200 needs a CUDA device: there is no CPU path")
    return torch.device("cuda", torch.cuda.current_device())


def _graph_of(data) -> CSRGraph:
    """data.norm_adj (scipy; raw COO with duplicates in ncl.py/directau.py, normalised CSR in selfcf.py) -> device CSR,
    cached on the data object so that several encoders over the same data share it."""
    g = getattr(data, "_gcf_graph", None)
    if g is None:
        adj = data.norm_adj
        g = adj if isinstance(adj, CSRGraph) else CSRGraph.from_scipy(adj, norm="none", device=_device())
        try:
            data._gcf_graph = g
        except AttributeError:
            pass
    return g


class _LGCNBase(JointEmbeddingDict):
    def __init__(self, data, emb_size: int, n_layers: int):
        super().__init__()
        self.data = data
        self.latent_size = emb_size
        self.layers = n_layers
        self.norm_adj = data.norm_adj
        init = nn.init.xavier_uniform_
        self.embedding_dict = nn.ParameterDict({
            "user_emb": nn.Parameter(init(torch.empty(data.user_num, emb_size))),
            "item_emb": nn.Parameter(init(torch.empty(data.item_num, emb_size))),
        })
        self.to(_device())  # also re-seats the two parameters on one allocation (JointEmbeddingDict._apply)
        self.sparse_norm_adj = _graph_of(data)  # CSRGraph in place of the torch.sparse COO tensor


class LGCNEncoder(_LGCNBase):
    """ncl.py:397-422 / directau.py:269-293."""

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor, List[torch.Tensor]]:
        emb = self.joint_table()
        final, layers = F_.propagate(self.sparse_norm_adj, emb, self.layers, mode="mean", return_layers=True)
        u = self.data.user_num
        return final[:u], final[u:], [emb, *layers]


class LGCN_Encoder(_LGCNBase):
    """selfcf.py:457-485."""

    def forward(self) -> Tuple[torch.Tensor, torch.Tensor]:
        final = F_.propagate(self.sparse_norm_adj, self.joint_table(), self.layers, mode="mean")
        u = self.data.user_num
        return final[:u], final[u:]


class SelfCF_HE(nn.Module):
    """selfcf.py:488-525.  u_target_his / i_target_his are plain tensor attributes (absent from state_dict), as in the
    reference; they live on the model's device."""

    def __init__(self, data, emb_size: int, momentum: float, n_layers: int):
        super().__init__()
        self.user_count = data.user_num
        self.item_count = data.item_num
        self.latent_size = emb_size
        self.momentum = momentum
        self.online_encoder = LGCN_Encoder(data, emb_size, n_layers)
        dev = _device()
        self.predictor = nn.Linear(self.latent_size, self.latent_size).to(dev)
        self.u_target_his = torch.randn((self.user_count, self.latent_size), requires_grad=False, device=dev)
        self.i_target_his = torch.randn((self.item_count, self.latent_size), requires_grad=False, device=dev)

    def forward(self, inputs: Dict[str, object]):
        u_online, i_online = self.online_encoder()
        dev = u_online.device
        users = F_._idx(inputs["user"], dev, "user")
        items = F_._idx(inputs["item"], dev, "item")
        u_rows = F_.gather_rows(u_online, users)   # gathered once; the scatter-add backward feeds the encoder
        i_rows = F_.gather_rows(i_online, items)
        with torch.no_grad():
            u_target = self.u_target_his[users] * self.momentum + u_rows.detach() * (1.0 - self.momentum)
            i_target = self.i_target_his[items] * self.momentum + i_rows.detach() * (1.0 - self.momentum)
            self.u_target_his[users, :] = u_rows.detach()
            self.i_target_his[items, :] = i_rows.detach()
        return self.predictor(u_rows), u_target, self.predictor(i_rows), i_target

    @torch.no_grad()
    def get_embedding(self):
        u_online, i_online = self.online_encoder.forward()
        return self.predictor(u_online), u_online, self.predictor(i_online), i_online

    def loss_fn(self, p, z):
        return 1 - TF.cosine_similarity(p, z.detach(), dim=-1).mean()

    def get_loss(self, output):
        u_online, u_target, i_online, i_target = output
        return self.loss_fn(u_online, i_target) / 2 + self.loss_fn(i_online, u_target) / 2


class DNNEncoder(nn.Module):
    """ssl4rec.py:162-196: two MLP towers over gathered embeddings (cuBLAS GEMMs, as in the reference) + InfoNCE between
    two dropout views on the tcgen05 kernel."""

    def __init__(self, data, emb_size: int, drop_rate: float, tau: float, n_layers: int):
        super().__init__()
        dev = _device()
        self.emb_size = emb_size
        self.tau = tau
        self.dropout = nn.Dropout(drop_rate)
        self.n_layers = n_layers
        init = nn.init.xavier_uniform_
        self.initial_user = nn.Parameter(init(torch.empty(data.user_num, emb_size, device=dev)))
        self.initial_item = nn.Parameter(init(torch.empty(data.item_num, emb_size, device=dev)))
        self.user_net = self.build_mlp(emb_size)
        self.item_net = self.build_mlp(emb_size)
        self.to(dev)

    def build_mlp(self, input_dim: int) -> nn.Sequential:
        layers = []
        hidden_dim = 1024
        for i in range(self.n_layers):
            out_dim = hidden_dim if i < self.n_layers - 1 else 128
            layers.append(nn.Linear(input_dim, out_dim))
            layers.append(nn.ReLU() if i < self.n_layers - 1 else nn.Tanh())
            input_dim = out_dim
        return nn.Sequential(*layers)

    def forward(self, u, i):
        return self.user_net(F_.gather_rows(self.initial_user, u)), self.item_net(F_.gather_rows(self.initial_item, i))

    def cal_cl_loss(self, i):
        emb = F_.gather_rows(self.initial_item, i)
        i1, i2 = self.dropout(emb), self.dropout(emb)
        return F_.info_nce(self.item_net(i1), self.item_net(i2), self.tau)


class EdgeRemoving:
    """gcl.py:18-25: keep each edge with probability 1 - pe."""

    def __init__(self, pe: float = 0.2):
        self.pe = pe

    def __call__(self, edge_index: torch.Tensor) -> torch.Tensor:
        keep_mask = torch.rand(edge_index.size(1), device=edge_index.device) >= self.pe
        return edge_index[:, keep_mask]


class GRACEModel(nn.Module):
    """gcl.py:38-64.  As in the reference, `encode` applies the Linear `convs` to the joint table and ignores the edge
    index it is given (the augmented graphs only enter through EdgeRemoving at the call site)."""

    def __init__(self, num_users: int, num_items: int, emb_size: int = 64, num_layers: int = 2, proj_dim: int = 64):
        super().__init__()
        self.user_emb = nn.Embedding(num_users, emb_size)
        self.item_emb = nn.Embedding(num_items, emb_size)
        self.convs = nn.ModuleList([nn.Linear(emb_size, emb_size) for _ in range(num_layers)])
        self.proj_head = nn.Sequential(nn.Linear(emb_size, proj_dim), nn.ReLU(), nn.Linear(proj_dim, proj_dim))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.xavier_uniform_(self.user_emb.weight)
        nn.init.xavier_uniform_(self.item_emb.weight)

    def encode(self, edge_index) -> torch.Tensor:
        x = torch.cat([self.user_emb.weight, self.item_emb.weight], dim=0)
        for conv in self.convs:
            x = conv(x)
        return x

    def project(self, x: torch.Tensor) -> torch.Tensor:
        return self.proj_head(x)

    def forward(self, edge_index1, edge_index2):
        return self.project(self.encode(edge_index1)), self.project(self.encode(edge_index2))
