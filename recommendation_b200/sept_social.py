"""Drop-in for the model part of univariate/sept_social.py (SEPT with social views).

    get_birectional_social_mat(S)                          sept_social.py:141-144   S o S (element-wise, as written there)
    get_social_related_views(social_mat, interaction_mat)  sept_social.py:361-368   [(S.S) o S + I, (Y.Y^T) o S + I], D^-1/2 . D^-1/2
    SEPTSocial                                             sept_social.py:333-420 and the loop body of 431-461
        .encoder / .social_encoder     sum over [E0, normalize(A E0), normalize(A normalize(A E0)), ...]
        .label_prediction / .sampling / .generate_pesudo_labels / .neighbor_discrimination / .iteration_losses

The two view matrices are masked sparse products evaluated on S's pattern only (gcf_spgemm_masked, never forming S.S or
Y.Y^T); every encoder layer is one SpMM launch with the row-L2-normalise epilogue fused in; the B_u x B_u denominators of the
neighbour-discrimination loss come from the tensor-core log-sum-exp kernel (no B_u x B_u matrix with autograd state) and the
pseudo-label top-K from the exact selection kernel of the evaluation path.  Dense B_u x B_u probabilities of label_prediction are
a cuBLAS GEMM + softmax exactly as in the reference (no gradient flows through them: only their top-K indices are used).

Note on the reference: its augmented branch (sept_social.py:425-427) calls `self.data.convert_to_laplacian_mat`, which the
Interaction class of that file does not define, so there `aug_mat` is always `norm_adj` in the epochs that run;
`iteration_losses(..., aug_adj=...)` accepts any operator (e.g. sept.GraphAugmentor.edge_dropout of the adjacency).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as TF

from . import _lib
from . import functional as F_
from . import motifs
from .encoders import _device
from .graph import CSRGraph

MAX_INSTANCES = 128   # gcf_masked_topn's list length limit


def _graph(m, dev) -> CSRGraph:
    return m if isinstance(m, CSRGraph) else CSRGraph.from_scipy(m, norm="none", device=dev)


def get_birectional_social_mat(social_mat) -> CSRGraph:
    """`social_mat.multiply(social_mat)` (sept_social.py:141-144): same pattern, squared values."""
    s = _graph(social_mat, _device())
    return s.with_values(s.vals * s.vals)


def get_social_related_views(social_mat, interaction_mat) -> List[CSRGraph]:
    """[social_matrix, sharing_matrix] of sept_social.py:361-368 as sym-normalised device operators."""
    dev = _device()
    s, y = _graph(social_mat, dev), _graph(interaction_mat, dev)
    n = s.n_rows
    if s.n_cols != n or y.n_rows != n:
        raise ValueError("get_social_related_views: social_mat must be [U, U] and interaction_mat [U, I]")
    eye = torch.arange(n, dtype=torch.int64, device=dev)
    diag = (eye, eye, torch.ones(n, dtype=torch.float32, device=dev))
    friends = motifs.to_coo(s, motifs.masked_product(s, s.transpose(), s))    # (S . S) o S
    sharing = motifs.to_coo(s, motifs.masked_product(y, y, s))                # (Y . Y^T) o S
    return [motifs.from_coo_sum([t, diag], n, n, dev, norm="sym") for t in (friends, sharing)]


class SEPTSocial(nn.Module):
    """state_dict keys: user_embeddings, item_embeddings (nn.Parameter, xavier_uniform_; sept_social.py:348-351)."""

    def __init__(self, data, social_mat, emb_size: int = 64, n_layers: int = 2, ss_rate: float = 0.005, ins_cnt: int = 10,
                 reg: float = 1e-4):
        """data: .user_num, .item_num, .norm_adj (scipy, the raw COO adjacency of sept_social.py:264-273),
        .interaction_mat (scipy [U, I]); social_mat: what `social_data.get_birectional_social_mat()` returns."""
        super().__init__()
        if not 1 <= ins_cnt <= MAX_INSTANCES:
            raise ValueError(f"ins_cnt must be in [1, {MAX_INSTANCES}]")
        dev = _device()
        self.data, self.n_layers, self.ss_rate, self.instance_cnt, self.reg = data, n_layers, ss_rate, ins_cnt, reg
        self.user_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(data.user_num, emb_size, device=dev)))
        self.item_embeddings = nn.Parameter(nn.init.xavier_uniform_(torch.empty(data.item_num, emb_size, device=dev)))
        self.bi_social_mat = _graph(social_mat, dev)
        self.norm_adj: Optional[CSRGraph] = None
        self.social_mat = self.sharing_mat = None
        self.aug_user_embeddings = None

    def build(self) -> None:
        """sept_social.py:387-392."""
        dev = _device()
        self.social_mat, self.sharing_mat = get_social_related_views(self.bi_social_mat, self.data.interaction_mat)
        self.norm_adj = _graph(self.data.norm_adj, dev)

    # ---- encoders (sept_social.py:370-385): the NORMALISED layer output feeds the next layer ----
    def _layers_sum(self, emb: torch.Tensor, adj: CSRGraph, n_layers: int) -> torch.Tensor:
        all_embs = [emb]
        for _ in range(n_layers):
            emb = F_.spmm_normalize(adj, emb)
            all_embs.append(emb)
        return torch.stack(all_embs, dim=0).sum(dim=0)

    def encoder(self, emb: torch.Tensor, adj: CSRGraph, n_layers: int) -> Tuple[torch.Tensor, torch.Tensor]:
        return torch.split(self._layers_sum(emb, adj, n_layers), [self.data.user_num, self.data.item_num], 0)

    def social_encoder(self, emb: torch.Tensor, adj: CSRGraph, n_layers: int) -> torch.Tensor:
        return self._layers_sum(emb, adj, n_layers)

    # ---- tri-training pieces (sept_social.py:394-420) ----
    def _batch_views(self, emb: torch.Tensor, u_idx: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        unique_u = torch.unique(u_idx)
        return F_.gather_rows(emb, unique_u), F_.gather_rows(self.aug_user_embeddings, unique_u)

    def label_prediction(self, emb: torch.Tensor, u_idx: torch.Tensor) -> torch.Tensor:
        e, a = self._batch_views(emb, u_idx)
        return TF.softmax(torch.matmul(TF.normalize(e), TF.normalize(a).T), dim=1)

    def sampling(self, logits: torch.Tensor) -> torch.Tensor:
        """torch.topk(logits, instance_cnt, dim=1).indices -- exact selection, ties broken by the lower column."""
        lib = _lib.load()
        logits = logits.detach().to(torch.float32).contiguous().clone()   # the kernel masks in place
        b, n = logits.shape
        k = min(self.instance_cnt, n)
        rows = torch.arange(b, dtype=torch.int64, device=logits.device)
        idx = torch.empty(b, k, dtype=torch.int64, device=logits.device)
        val = torch.empty(b, k, dtype=torch.float32, device=logits.device)
        _lib.check(lib.gcf_masked_topn(_lib.ptr(logits), logits.stride(0), b, n, _lib.ptr(rows), None, None, -1e8, k,
                                       _lib.ptr(idx), _lib.ptr(val), _lib.current_stream()), "gcf_masked_topn")
        return idx

    def generate_pesudo_labels(self, prob1: torch.Tensor, prob2: torch.Tensor) -> torch.Tensor:
        return self.sampling((prob1 + prob2) / 2)

    def neighbor_discrimination(self, positive: torch.Tensor, emb: torch.Tensor, u_idx: torch.Tensor) -> torch.Tensor:
        """-sum_i log( sum_{k in positive_i} exp(e_i a_k / 0.1) / sum_j exp(e_i a_j / 0.1) ) over the unique batch users."""
        e, a = self._batch_views(emb, u_idx)
        row_lse, _, _ = F_.infonce_stats(e, a, 0.1, cos=True)              # log sum_j exp(cos / 0.1): tensor cores, no B x B matrix
        e_hat, a_hat = TF.normalize(e), TF.normalize(a)
        pos = torch.sum(e_hat.unsqueeze(1) * a_hat[positive], dim=2)       # [B_u, K]
        return -torch.sum(torch.logsumexp(pos / 0.1, dim=1) - row_lse)

    # ---- one iteration of the training loop (sept_social.py:431-461) ----
    def iteration_losses(self, user_idx, pos_idx, neg_idx, *, ssl: bool = True, aug_adj: Optional[CSRGraph] = None, labels=None):
        """Returns (rec_loss, neighbor_dis_loss | None, total_loss); sets the rec_/aug_/view embedding attributes like the
        reference loop does.  labels: optional (f_pos, sh_pos, r_pos) index tensors replacing this run's pseudo-labels
        (parity runs inject the reference's; training leaves it None)."""
        dev = self.user_embeddings.device
        to_idx = lambda t: torch.as_tensor(t, dtype=torch.int64, device=dev)
        user_idx, pos_idx, neg_idx = to_idx(user_idx), to_idx(pos_idx), to_idx(neg_idx)
        if self.norm_adj is None:
            self.build()
        ego = torch.cat([self.user_embeddings, self.item_embeddings], dim=0)
        self.rec_user_embeddings, self.rec_item_embeddings = self.encoder(ego, self.norm_adj, self.n_layers)
        if aug_adj is None:
            self.aug_user_embeddings, self.aug_item_embeddings = self.rec_user_embeddings, self.rec_item_embeddings
        else:
            self.aug_user_embeddings, self.aug_item_embeddings = self.encoder(ego, aug_adj, self.n_layers)
        self.sharing_view_embeddings = self.social_encoder(self.user_embeddings, self.sharing_mat, self.n_layers)
        self.friend_view_embeddings = self.social_encoder(self.user_embeddings, self.social_mat, self.n_layers)
        rec_loss = F_.bpr_loss_gather(self.rec_user_embeddings, self.rec_item_embeddings, user_idx, pos_idx, neg_idx,
                                      variant="softplus")                 # -mean(logsigmoid(pos - neg)), sept_social.py:37-41
        rec_loss = rec_loss + self.reg * (self.user_embeddings.norm(2).pow(2) + self.item_embeddings.norm(2).pow(2))
        if not ssl:
            return rec_loss, None, rec_loss
        with torch.no_grad():
            social_p = self.label_prediction(self.friend_view_embeddings, user_idx)
            sharing_p = self.label_prediction(self.sharing_view_embeddings, user_idx)
            rec_p = self.label_prediction(self.rec_user_embeddings, user_idx)
            f_pos = self.generate_pesudo_labels(sharing_p, rec_p)
            sh_pos = self.generate_pesudo_labels(social_p, rec_p)
            r_pos = self.generate_pesudo_labels(social_p, sharing_p)
        self.last_predictions = (social_p, sharing_p, rec_p)
        self.last_labels = (f_pos, sh_pos, r_pos)
        if labels is not None:
            f_pos, sh_pos, r_pos = (torch.as_tensor(t, dtype=torch.int64, device=dev) for t in labels)
        nd = self.neighbor_discrimination(f_pos, self.friend_view_embeddings, user_idx)
        nd = nd + self.neighbor_discrimination(sh_pos, self.sharing_view_embeddings, user_idx)
        nd = nd + self.neighbor_discrimination(r_pos, self.rec_user_embeddings, user_idx)
        self.last_nd = nd
        return rec_loss, nd, rec_loss + self.ss_rate * nd
