"""Drop-in for the reference's lightgcn.py model and training step, on the libgcf kernels.

Reference interface kept (lightgcn.py:12-27, 77-120):
    LightGCN(num_users, num_items, embedding_dim=64, num_layers=3)
        .user_embedding / .item_embedding : nn.Embedding   (state_dict keys user_embedding.weight, item_embedding.weight)
        .forward(edge_index[2, 2E] int64) -> (user_emb[U, d], item_emb[I, d])     x = sum_{k=0..K} E(k)
    build_edge_index(users, items, num_users)             the edge_index of load_data (lightgcn.py:36-39)
    bpr_step_loss(...)                                    lightgcn.py:95-118 (BPR + reg on the gathered rows)
    train_step(...)                                       one iteration of the epoch loop, lightgcn.py:83-120

B200-first differences (results identical within fp32 tolerance):
  * gcn_norm + gather/scatter message passing recomputed K times per forward in PyG becomes ONE cached CSR
    build (integer kernels) and K fused SpMM launches; the layer sum is folded into the last SpMM epilogue.
  * the [E, d] gathered tensors and their index_put backward become one fused gather+BPR kernel each way.
  * both embedding tables live in one [U+I, d] allocation so torch.cat / its backward disappear.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import functional as F_
from .graph import CSRGraph
from .tables import JoinTables as _JoinTables, join_parameters


def build_edge_index(users: torch.Tensor, items: torch.Tensor, num_users: int) -> torch.Tensor:
    """[[u | i+U], [i+U | u]] int64 -- load_data's edge_index (lightgcn.py:36-39), built on the GPU when the
    inputs are CUDA tensors."""
    users = users.to(torch.int64)
    items = items.to(torch.int64)
    if users.is_cuda:
        lib = _lib.load()
        e = users.numel()
        out = torch.empty(2, 2 * e, dtype=torch.int64, device=users.device)
        _lib.check(lib.gcf_bipartite_edge_index(_lib.ptr(users.contiguous()), _lib.ptr(items.contiguous()), e, num_users,
                                                _lib.ptr(out[0]), _lib.ptr(out[1]), _lib.current_stream()),
                   "gcf_bipartite_edge_index")
        return out
    return torch.stack([torch.cat([users, items + num_users]), torch.cat([items + num_users, users])])


class LGConv(nn.Module):
    """Parameter-less placeholder keeping `model.convs` shaped like the reference (lightgcn.py:17)."""

    def forward(self, x: torch.Tensor, graph: CSRGraph) -> torch.Tensor:
        return F_.spmm(graph, x)


class LightGCN(nn.Module):
    def __init__(self, num_users: int, num_items: int, embedding_dim: int = 64, num_layers: int = 3):
        super().__init__()
        self.user_embedding = nn.Embedding(num_users, embedding_dim)
        self.item_embedding = nn.Embedding(num_items, embedding_dim)
        self.convs = nn.ModuleList([LGConv() for _ in range(num_layers)])
        nn.init.xavier_uniform_(self.user_embedding.weight)
        nn.init.xavier_uniform_(self.item_embedding.weight)
        self._graph_cache: Dict[Tuple, CSRGraph] = {}
        self._join()

    # -- one allocation for both tables ------------------------------------------------------
    def _join(self) -> None:
        self._table = join_parameters(self.user_embedding.weight, self.item_embedding.weight)

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._join()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._join()
        return out

    @property
    def table(self) -> torch.Tensor:
        """[U+I, d] storage shared by user_embedding.weight and item_embedding.weight."""
        return self._table

    # -- graph cache -------------------------------------------------------------------------
    def graph_for(self, edge_index: torch.Tensor) -> CSRGraph:
        """The reference re-normalises inside every LGConv call (K times per forward); the CSR is built once
        per distinct edge_index tensor and kept."""
        key = (edge_index.data_ptr(), tuple(edge_index.shape), edge_index._version, str(edge_index.device))
        g = self._graph_cache.get(key)
        if g is None:
            dev = self.user_embedding.weight.device
            n = self.user_embedding.num_embeddings + self.item_embedding.num_embeddings
            g = CSRGraph.from_edge_index(edge_index.to(dev), n, norm="sym")
            self._graph_cache = {key: g}
            self._edge_index_ref = edge_index  # keep the keyed tensor alive so data_ptr is not recycled
        return g

    def forward(self, edge_index) -> Tuple[torch.Tensor, torch.Tensor]:
        graph = edge_index if isinstance(edge_index, CSRGraph) else self.graph_for(edge_index)
        x0 = _JoinTables.apply(self.user_embedding.weight, self.item_embedding.weight)
        x = F_.propagate(graph, x0, len(self.convs), mode="sum")
        u = self.user_embedding.num_embeddings
        return x[:u], x[u:]


def bpr_step_loss(user_emb: torch.Tensor, item_emb: torch.Tensor, pos_u, pos_i, neg_i, reg_weight: float) -> torch.Tensor:
    """loss = mean(-log sigmoid(pos - neg)) + reg_weight * (|u_vecs|^2 + |pos_vecs|^2)   (lightgcn.py:95-118);
    neg_i may be [E] or [E, n_neg] (negative score = mean over the n_neg samples)."""
    return F_.bpr_loss_gather(user_emb, item_emb, pos_u, pos_i, neg_i, variant="softplus", reduction="mean",
                              reg_u=reg_weight, reg_p=reg_weight, reg_n=0.0)


def train_step(model: LightGCN, optimizer, edge_index, pos_u, pos_i, num_items: int, config: dict, *,
               neg_i: Optional[torch.Tensor] = None, seed: int = 0, step: int = 0) -> torch.Tensor:
    """One iteration of the reference epoch loop (lightgcn.py:83-120) through the public model API.
    Negatives: given, or drawn on the device with the Philox sampler (torch.randint semantics: no rejection)."""
    if config.get("loss_type", "bpr") != "bpr":
        raise NotImplementedError("loss_type='bce' materialises a dense [E, I] matrix (lightgcn.py:109-113); out of scope")
    optimizer.zero_grad(set_to_none=True)
    user_emb, item_emb = model(edge_index)
    dev = user_emb.device
    if neg_i is None:
        n = pos_u.numel() if torch.is_tensor(pos_u) else len(pos_u)
        neg_i = F_.sample_negatives(n, num_items, seed=seed, offset=step, n_negs=int(config.get("n_neg", 1)), device=dev)
    loss = bpr_step_loss(user_emb, item_emb, pos_u, pos_i, neg_i, float(config.get("reg_weight", 1e-4)))
    loss.backward()
    optimizer.step()
    return loss


class FusedLightGCNTrainer:
    """The same step as `train_step`, issued as a fixed sequence of libgcf launches without autograd:

        propagate_fwd (K SpMM, sum folded into the last) -> Philox negatives -> fused gather+BPR forward
        -> zero g_final -> fused BPR backward (red.add) -> propagate_bwd (K SpMM) -> fused Adam

    Used by bench.py for the device-resident `value`; `train_step` above is the reference-facing path.
    Triples are the training interactions (full batch, lightgcn.py:86-88).
    """

    def __init__(self, graph: CSRGraph, n_users: int, n_items: int, table: torch.Tensor, pos_u: torch.Tensor,
                 pos_i: torch.Tensor, *, n_layers: int = 3, lr: float = 0.01, reg_weight: float = 1e-4,
                 weight_decay: float = 0.0, n_neg: int = 1, seed: int = 0, sort_triples: bool = True,
                 fused_bpr: bool = True, fused_adam: bool = True):
        self.lib = _lib.load()
        self.fused_bpr = fused_bpr
        self.fused_adam = fused_adam
        self.order = None
        self.graph, self.n_users, self.n_items = graph, n_users, n_items
        self.table = table
        self.n, self.d = table.shape
        dev = table.device
        pos_u = pos_u.to(dev, torch.int64)
        pos_i = pos_i.to(dev, torch.int64)
        if sort_triples and pos_u.numel() > 1:
            # full batch: the order of the interactions is free (the loss is a mean over all of them, lightgcn.py:86-88).
            # User-major order lets the fused kernel keep the user row and its gradient in registers over a run.
            order = torch.argsort(pos_u * n_items + pos_i)
            pos_u, pos_i = pos_u[order], pos_i[order]
            self.order = order  # externally supplied negatives (parity runs) are permuted the same way
        self.pos_u = pos_u.contiguous()
        self.pos_i = pos_i.contiguous()
        self.n_triples = self.pos_u.numel()
        self.k, self.lr, self.reg, self.wd, self.n_neg, self.seed = n_layers, lr, reg_weight, weight_decay, n_neg, seed
        self.layers = [torch.empty_like(table) for _ in range(n_layers - 1)] + [None]
        self.final = torch.empty_like(table)
        self.g_final = torch.empty_like(table)
        self.ping = torch.empty_like(table) if n_layers > 1 else None
        self.pong = torch.empty_like(table) if n_layers > 1 else None
        self.g_x0 = None if fused_adam else torch.empty_like(table)  # fused Adam consumes G(0) inside the last SpMM
        self.exp_avg = torch.zeros_like(table)
        self.exp_avg_sq = torch.zeros_like(table)
        self.neg = torch.empty(self.n_triples * n_neg, dtype=torch.int64, device=dev)
        self.coef = torch.empty(max(self.n_triples, 1), dtype=torch.float32, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.bpr_ws_bytes = self.lib.gcf_bpr_workspace_bytes(self.n_triples)
        self.bpr_ws = torch.empty(self.bpr_ws_bytes, dtype=torch.uint8, device=dev)
        self.ws, self.ws_bytes = graph.workspace(self.d)
        self.step_count = 0
        # libgcf kernels only: K fwd + K bwd SpMM, sampler, bpr (fused fwd+bwd | fwd, bwd) + reduce, adam
        self.launches_per_step = 2 * n_layers + (4 if fused_bpr else 5) - (1 if fused_adam else 0)

    def step(self, neg_i: Optional[torch.Tensor] = None, marks: Optional[list] = None) -> torch.Tensor:
        """One optimisation step.  `marks` (optional list) receives (start_event, end_event, n_spmm_launches)
        tuples bracketing the SpMM launches, so a caller can time the dominant kernel inside the step."""
        lib, st = self.lib, _lib.current_stream()
        g, d, u = self.graph, self.d, self.n_users
        self.step_count += 1
        if marks is not None:
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
        _lib.check(lib.gcf_propagate_fwd(g.struct_ref(), d, self.k, _lib.ptr(self.table), _lib.ptr_array(self.layers),
                                         _lib.ptr(self.final), 1.0, _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_fwd")
        if marks is not None:
            e1.record()
        if neg_i is None:
            _lib.check(lib.gcf_sample_negatives(self.seed, self.step_count, None, self.n_triples, self.n_neg, self.n_items,
                                                None, None, 1, _lib.ptr(self.neg), st), "gcf_sample_negatives")
            neg = self.neg
        else:
            neg = neg_i.reshape(self.n_triples, -1)
            if self.order is not None:
                neg = neg[self.order]
            neg = neg.reshape(-1).contiguous()
        user_emb, item_emb = self.final[:u], self.final[u:]
        if self.fused_bpr:
            self.g_final.zero_()
            _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(user_emb), d, _lib.ptr(item_emb), d, d, _lib.ptr(self.pos_u), _lib.ptr(self.pos_i),
                                           _lib.ptr(neg), self.n_triples, self.n_neg, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_MEAN,
                                           self.reg, self.reg, 0.0, 1.0, _lib.ptr(self.loss), None,
                                           _lib.ptr(self.g_final[:u]), d, _lib.ptr(self.g_final[u:]), d, _lib.ptr(self.bpr_ws),
                                           self.bpr_ws_bytes, st), "gcf_bpr_fwd_bwd")
        else:
            _lib.check(lib.gcf_bpr_fwd(_lib.ptr(user_emb), d, _lib.ptr(item_emb), d, d, _lib.ptr(self.pos_u), _lib.ptr(self.pos_i),
                                       _lib.ptr(neg), self.n_triples, self.n_neg, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_MEAN,
                                       self.reg, self.reg, 0.0, _lib.ptr(self.loss), _lib.ptr(self.coef), _lib.ptr(self.bpr_ws),
                                       self.bpr_ws_bytes, st), "gcf_bpr_fwd")
            self.g_final.zero_()
            _lib.check(lib.gcf_bpr_bwd(_lib.ptr(user_emb), d, _lib.ptr(item_emb), d, d, _lib.ptr(self.pos_u), _lib.ptr(self.pos_i),
                                       _lib.ptr(neg), self.n_triples, self.n_neg, _lib.ptr(self.coef), None, self.reg, self.reg, 0.0,
                                       _lib.ptr(self.g_final[:u]), d, _lib.ptr(self.g_final[u:]), d, st), "gcf_bpr_bwd")
        gt = g.transpose()
        if marks is not None:
            e2.record()
        if self.fused_adam:
            _lib.check(lib.gcf_propagate_bwd_adam(gt.struct_ref(), d, self.k, _lib.ptr(self.g_final), None, 1.0, _lib.ptr(self.ping),
                                                  _lib.ptr(self.pong), None, _lib.ptr(self.table), _lib.ptr(self.exp_avg),
                                                  _lib.ptr(self.exp_avg_sq), self.lr, 0.9, 0.999, 1e-8, self.wd, 0, self.step_count,
                                                  _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_bwd_adam")
        else:
            _lib.check(lib.gcf_propagate_bwd(gt.struct_ref(), d, self.k, _lib.ptr(self.g_final), None, 1.0, _lib.ptr(self.ping),
                                             _lib.ptr(self.pong), _lib.ptr(self.g_x0), _lib.ptr(self.ws), self.ws_bytes, st),
                       "gcf_propagate_bwd")
        if marks is not None:
            e3.record()
            marks.append((e0, e1, self.k))
            marks.append((e2, e3, self.k))
        if not self.fused_adam:
            _lib.check(lib.gcf_adam_step(_lib.ptr(self.table), _lib.ptr(self.g_x0), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                         self.table.numel(), self.lr, 0.9, 0.999, 1e-8, self.wd, 0, self.step_count, st),
                       "gcf_adam_step")
        return self.loss


# =========================================================================================
# The rest of lightgcn.py's script surface: load_data, evaluate, train_model (lightgcn.py:29-124)
# =========================================================================================
def load_data(train_path: str, test_path: str, device=None):
    """`user item rating` text files -> (edge_index [2, 2E] int64 on the GPU, train_df, test_df, num_users, num_items),
    the return signature of lightgcn.py:29-40."""
    import pandas as pd

    train_df = pd.read_csv(train_path, sep=" ", names=["user", "item", "rating"])
    test_df = pd.read_csv(test_path, sep=" ", names=["user", "item", "rating"])
    num_users = int(max(train_df["user"].max(), test_df["user"].max()) + 1)
    num_items = int(max(train_df["item"].max(), test_df["item"].max()) + 1)
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    users = torch.from_numpy(train_df["user"].values.astype("int64")).to(dev)
    items = torch.from_numpy(train_df["item"].values.astype("int64")).to(dev)
    return build_edge_index(users, items, num_users), train_df, test_df, num_users, num_items


def evaluate(user_emb: torch.Tensor, item_emb: torch.Tensor, test_df, train_pos=None, k_list=(10,), *, train_df=None):
    """lightgcn.py:48-74 for all test users at once: HR = share of users with at least one hit, P = hits / k, R = hits /
    |test items|, NDCG = the un-normalised DCG the reference accumulates, each averaged over the test users.
    `train_pos` ({user: set(items)}, as get_user_positive_items returns) or `train_df` names the items to exclude."""
    from . import evaluation

    dev = user_emb.device
    n_users, n_items = user_emb.shape[0], item_emb.shape[0]
    if train_df is not None:
        tr_u, tr_i = train_df["user"].values, train_df["item"].values
    elif train_pos:
        tr_u = [u for u, its in train_pos.items() for _ in its]
        tr_i = [i for its in train_pos.values() for i in its]
    else:
        tr_u, tr_i = [], []
    as_t = lambda a: torch.tensor(np.asarray(a, dtype=np.int64)).to(dev)   # (copy: pandas hands out read-only views)
    train_csr = evaluation.positives_csr(as_t(tr_u), as_t(tr_i), n_users, n_items) if len(tr_u) else None
    te_u, te_i = as_t(test_df["user"].values), as_t(test_df["item"].values)
    test_csr = evaluation.positives_csr(te_u, te_i, n_users, n_items)
    query = torch.unique(te_u)
    ks = sorted(int(k) for k in k_list)
    idx, _ = evaluation.recommend_topn(user_emb, item_emb, query, ks[-1], train_pos=train_csr)
    hits, dcg = evaluation.hits_and_dcg(idx, query, test_csr, ks)
    n_test = (test_csr[0][query + 1] - test_csr[0][query]).to(torch.float64)
    out = {}
    for c, k in enumerate(ks):
        h = hits[:, c].to(torch.float64)
        out[k] = {"HR": float((h > 0).double().mean()), "P": float((h / k).mean()), "R": float((h / n_test).mean()),
                  "NDCG": float(dcg[:, c].double().mean())}
    return out


def train_model(config: dict, train_path: str = "./data/train.txt", test_path: str = "./data/test.txt", epochs: int = 30,
                seed: int = 0):
    """lightgcn.py:77-124: 30 full-batch epochs, then evaluation at k = 10."""
    edge_index, train_df, test_df, num_users, num_items = load_data(train_path, test_path)
    dev = edge_index.device
    model = LightGCN(num_users, num_items, embedding_dim=config["embedding_dim"], num_layers=config["num_layers"]).to(dev)
    optimizer = getattr(torch.optim, config["optimizer"])(model.parameters(), lr=config["lr"], weight_decay=config["weight_decay"])
    pos_u = torch.from_numpy(train_df["user"].values.astype("int64")).to(dev)
    pos_i = torch.from_numpy(train_df["item"].values.astype("int64")).to(dev)
    model.train()
    for epoch in range(epochs):
        train_step(model, optimizer, edge_index, pos_u, pos_i, num_items, config, seed=seed, step=epoch)
    model.eval()
    with torch.no_grad():
        user_emb, item_emb = model(edge_index)
        return evaluate(user_emb, item_emb, test_df, train_df=train_df, k_list=[10])
