"""Oracle: the loss functions of the hot path, restated for torch-CPU tensors.  Test infrastructure only.

Written from the formulas in SURVEY.md 8(a) "Exact formulas" and pinned against the reference's own functions
through tests/golden (tests/test_oracle_golden.py).  All functions are differentiable with torch autograd, which is
how the parity tests obtain reference gradients.  `dtype` of the inputs decides the precision (fp32 like the
reference, or fp64 for a 'true value' check).
"""
from __future__ import annotations

import torch


def _unit_rows(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    # F.normalize(x, dim=1): x / max(|x|_2, eps)
    return x / x.norm(dim=1, keepdim=True).clamp_min(eps)


def bpr_log_eps_sigmoid(u, p, n, eps: float = 1e-5, reduction: str = "mean"):
    """ncl.py:116-120, mhcn.py:35-39: mean(-log(1e-5 + sigmoid(<u,p> - <u,n>)))."""
    x = (u * p).sum(1) - (u * n).sum(1)
    l = -torch.log(eps + torch.sigmoid(x))
    return l.mean() if reduction == "mean" else l.sum()


def bpr_lightgcn(user_emb, item_emb, pos_u, pos_i, neg_i, reg_weight: float):
    """lightgcn.py:95-118: mean(-log sigmoid(pos - neg)) + reg * (|u_vecs|^2 + |pos_vecs|^2); neg_i [E] or [E, n_neg]."""
    uv, pv = user_emb[pos_u], item_emb[pos_i]
    nv = item_emb[neg_i]
    pos = (uv * pv).sum(-1)
    if neg_i.dim() == 1:
        neg = (uv * nv).sum(-1)
    else:
        neg = (uv.unsqueeze(1) * nv).sum(-1).mean(1)
    loss = -torch.log(torch.sigmoid(pos - neg)).mean()
    return loss + reg_weight * (uv.norm(2).pow(2) + pv.norm(2).pow(2))


def bpr_gcl(u_e, p_e, n_e, reg_weight: float):
    """gcl.py:219-223: -logsigmoid(pos - neg).mean() + reg * (|u|^2 + |p|^2 + |n|^2) / B."""
    x = (u_e * p_e).sum(1) - (u_e * n_e).sum(1)
    reg = (u_e.norm(2).pow(2) + p_e.norm(2).pow(2) + n_e.norm(2).pow(2)) / u_e.shape[0]
    return -torch.nn.functional.logsigmoid(x).mean() + reg_weight * reg


def l2_reg(reg: float, *xs):
    """ncl.py:122-123 (= directau.py:35, ssl4rec.py:16): reg * sum_x |x|_F / rows(x)   (norm NOT squared)."""
    return reg * sum(x.norm(p=2) / x.shape[0] for x in xs)


def info_nce(v1, v2, tau: float, cos: bool = True):
    """ncl.py:125-130 = ssl4rec.py:19-23: -mean_i (s_ii - logsumexp_j s_ij), s = v1^ v2^T / tau."""
    if cos:
        v1, v2 = _unit_rows(v1), _unit_rows(v2)
    s = v1 @ v2.T / tau
    return -(s.diagonal() - torch.logsumexp(s, dim=1)).mean()


def ssl_layer_loss(context, initial, user_idx, item_idx, n_users: int, tau: float, ssl_reg: float, alpha: float):
    """ncl.py:358-367: structure-contrastive loss with the denominator over ALL users / ALL items (sum over batch)."""
    cu, ci = context[:n_users], context[n_users:]
    iu, ii = initial[:n_users], initial[n_users:]

    def side(c, z, idx):
        cb, zb = _unit_rows(c[idx]), _unit_rows(z[idx])
        pos = (cb * zb).sum(1) / tau
        all_ = cb @ _unit_rows(z).T / tau
        return (torch.logsumexp(all_, dim=1) - pos).sum()

    return ssl_reg * (side(cu, iu, user_idx) + alpha * side(ci, ii, item_idx))


def proto_nce(initial, user_idx, item_idx, n_users: int, user_centroids, user_2cluster, item_centroids, item_2cluster,
              tau: float, proto_reg: float, batch_size_conf: int):
    """ncl.py:369-375: proto_reg * B_conf * (InfoNCE(z_u, C_u[a_u]) + InfoNCE(z_i, C_i[a_i]))."""
    ue, ie = initial[:n_users], initial[n_users:]
    lu = info_nce(ue[user_idx], user_centroids[user_2cluster[user_idx]], tau) * batch_size_conf
    li = info_nce(ie[item_idx], item_centroids[item_2cluster[item_idx]], tau) * batch_size_conf
    return proto_reg * (lu + li)


def batch_softmax(u, i, tau: float):
    """ssl4rec.py:25-30: -mean log( exp(s_ii) / sum_j exp(s_ij) + 1e-6 )."""
    u, i = _unit_rows(u), _unit_rows(i)
    s = u @ i.T / tau
    p = torch.exp(s.diagonal() - torch.logsumexp(s, dim=1))
    return -torch.log(p + 1e-6).mean()


def info_nce_symmetric(z1, z2, temp: float = 0.2):
    """gcl.py:28-35: (CE(S, arange) + CE(S^T, arange)) / 2."""
    z1, z2 = _unit_rows(z1), _unit_rows(z2)
    s = z1 @ z2.T / temp
    d = s.diagonal()
    return 0.5 * ((torch.logsumexp(s, 1) - d).mean() + (torch.logsumexp(s, 0) - d).mean())


def alignment(x, y):
    """directau.py:245: mean |x^ - y^|^2."""
    return (_unit_rows(x) - _unit_rows(y)).pow(2).sum(1).mean()


def uniformity(x, t: float = 2.0):
    """directau.py:248-251: log(mean_{i<j} exp(-t |x^_i - x^_j|^2) + 1e-8); 0 when there are no pairs."""
    x = _unit_rows(x)
    b = x.shape[0]
    if b < 2:
        return x.new_zeros(())
    iu = torch.triu_indices(b, b, offset=1)
    sq = (x[iu[0]] - x[iu[1]]).pow(2).sum(1)
    return ((-t * sq).exp().mean() + 1e-8).log()


def directau_loss(u, i, gamma: float):
    """directau.py:240-243: align + gamma * (unif(u) + unif(i)) / 2."""
    return alignment(u, i) + gamma * (uniformity(u) + uniformity(i)) / 2


def selfcf_loss(p_u, t_u, p_i, t_i):
    """selfcf.py:518-525: (1 - mean cos(p_u, t_i))/2 + (1 - mean cos(p_i, t_u))/2, targets detached."""
    def one(p, z):
        return 1 - torch.nn.functional.cosine_similarity(p, z.detach(), dim=-1).mean()
    return one(p_u, t_i) / 2 + one(p_i, t_u) / 2
