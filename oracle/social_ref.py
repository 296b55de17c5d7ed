"""Oracle: CPU restatement (torch fp64, dense adjacency) of the social-graph models' forward passes.  Test infrastructure only.

  mhcn_forward     univariate/mhcn.py:404-505 (self_gating, channel_attention, forward, hierarchical_self_supervision)
  diffnet_forward  univariate/diffnet.py:1124-1132
  sept_social_views / sept_social_iteration   univariate/sept_social.py:361-420 and the loop body of 431-461
Pinned against tests/golden/{mhcn_model,diffnet_model,sept_social}.npz, which the reference's own classes produced.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gate(em, w, b):
    return em * torch.sigmoid(em @ w + b)


def _attention(att, att_mat, *chs):
    w = torch.stack([(att * (e @ att_mat)).sum(1) for e in chs])
    score = F.softmax(w, dim=0)
    return sum(score[i].view(-1, 1) * chs[i] for i in range(len(chs)))


def hierarchical_self_supervision(em, adj, perms):
    """mhcn.py:480-505 with the three row permutations given explicitly."""
    score = lambda a, b: (a * b).sum(1)
    edge = adj @ em
    pos, neg1, neg2 = score(em, edge), score(em[perms[0]], edge), score(edge[perms[1]], em)
    local = torch.sum(-torch.log(torch.sigmoid(pos - neg1)) - torch.log(torch.sigmoid(neg1 - neg2)))
    graph = edge.mean(0, keepdim=True)
    glob = torch.sum(-torch.log(torch.sigmoid(score(edge, graph.expand_as(edge)) - score(edge[perms[2]], graph.expand_as(edge)))))
    return glob + local


def mhcn_forward(p, Hs, Hj, Hp, R, n_layers, ss_rate, u_idx, v_idx, neg_idx, perms):
    """p: dict of parameters keyed like the reference's state_dict; Hs/Hj/Hp/R dense fp64 matrices."""
    ue = p["user_embeddings"]
    c1, c2, c3, simple = (_gate(ue, p[f"gating_weights.{k}"], p[f"gating_bias.{k}"]) for k in (1, 2, 3, 4))
    a1, a2, a3, asimple = [c1], [c2], [c3], [simple]
    item = p["item_embeddings"]
    ai = [item]
    for _ in range(n_layers):
        mixed = _attention(p["attention"], p["attention_mat"], c1, c2, c3) + simple / 2
        c1 = Hs @ c1; a1.append(F.normalize(c1, dim=1))
        c2 = Hj @ c2; a2.append(F.normalize(c2, dim=1))
        c3 = Hp @ c3; a3.append(F.normalize(c3, dim=1))
        new_item = R.T @ mixed; ai.append(F.normalize(new_item, dim=1))
        simple = R @ item; asimple.append(F.normalize(simple, dim=1))
        item = new_item
    c1, c2, c3, simple, final_item = (torch.stack(a).sum(0) for a in (a1, a2, a3, asimple, ai))
    final_user = _attention(p["attention"], p["attention_mat"], c1, c2, c3) + simple / 2
    ss = 0
    for k, adj in ((1, Hs), (2, Hj), (3, Hp)):
        ss = ss + hierarchical_self_supervision(_gate(final_user, p[f"sgating_weights.{k}"], p[f"sgating_bias.{k}"]), adj, perms[3 * (k - 1):3 * k])
    return final_user[u_idx], final_item[v_idx], final_item[neg_idx], ss_rate * ss, final_user, final_item


def diffnet_forward(user_w, item_w, weights, S, A):
    user = user_w
    for w in weights:
        user = torch.relu(torch.cat([S @ user, user], dim=1) @ w)
    return user + A @ item_w


# ------------------------------------------------------------------------------------------ sept_social.py
def sept_social_views(S, Y):
    """[social_matrix, sharing_matrix] of sept_social.py:361-368 (scipy): (S.S) o S + I and (Y.Y^T) o S + I, each
    normalised with D^-1/2 . D^-1/2 on its ROW sums (normalize_graph_mat, sept_social.py:86-101)."""
    import numpy as np
    import scipy.sparse as sp

    def sym_norm(m):
        m = sp.csr_matrix(m, dtype=np.float32)
        rs = np.asarray(m.sum(1), dtype=np.float32).ravel()
        with np.errstate(divide="ignore"):
            dinv = np.power(rs, -0.5).astype(np.float32)
        dinv[np.isinf(dinv)] = 0.0
        return sp.diags(dinv).dot(m).dot(sp.diags(dinv)).tocsr().astype(np.float32)

    S = sp.csr_matrix(S, dtype=np.float32)
    Y = sp.csr_matrix(Y, dtype=np.float32)
    eye = sp.eye(S.shape[0], dtype=np.float32)
    return [sym_norm((S @ S).multiply(S) + eye), sym_norm((Y @ Y.T).multiply(S) + eye)]


def _sept_layers_sum(emb, adj, n_layers):
    """sept_social.py:370-385: the normalised layer output is what the next layer propagates."""
    out = [emb]
    for _ in range(n_layers):
        emb = F.normalize(adj @ emb)
        out.append(emb)
    return torch.stack(out, 0).sum(0)


def sept_social_iteration(user_w, item_w, adj, social, sharing, n_layers, ss_rate, ins_cnt, reg, u_idx, p_idx, n_idx, labels=None):
    """One iteration of sept_social.py:431-461 with aug_mat = norm_adj (dense fp64 operators).  labels: optional (f_pos, sh_pos,
    r_pos) to use instead of the top-K of this run.  Returns a dict of every intermediate the fixture stores."""
    nu = user_w.shape[0]
    ego = torch.cat([user_w, item_w], 0)
    rec = _sept_layers_sum(ego, adj, n_layers)
    rec_u, rec_i = rec[:nu], rec[nu:]
    aug_u = rec_u
    sharing_v, friend_v = _sept_layers_sum(user_w, sharing, n_layers), _sept_layers_sum(user_w, social, n_layers)
    x = (rec_u[u_idx] * rec_i[p_idx]).sum(1) - (rec_u[u_idx] * rec_i[n_idx]).sum(1)
    rec_loss = -F.logsigmoid(x).mean() + reg * (user_w.norm(2).pow(2) + item_w.norm(2).pow(2))
    uniq = torch.unique(u_idx)

    def predict(emb):                                       # label_prediction, :394-399
        return F.softmax(F.normalize(emb[uniq]) @ F.normalize(aug_u[uniq]).T, dim=1)

    def discriminate(positive, emb):                        # neighbor_discrimination, :408-420
        e, a = F.normalize(emb[uniq]), F.normalize(aug_u[uniq])
        pos = (e.unsqueeze(1) * a[positive]).sum(2)
        return -torch.sum(torch.log(torch.exp(pos / 0.1).sum(1) / torch.exp(e @ a.T / 0.1).sum(1)))

    soc_p, sha_p, rec_p = predict(friend_v), predict(sharing_v), predict(rec_u)
    topk = lambda a, b: torch.topk((a + b) / 2, ins_cnt, dim=1).indices   # generate_pesudo_labels, :401-406
    f_pos, sh_pos, r_pos = labels if labels is not None else (topk(sha_p, rec_p), topk(soc_p, rec_p), topk(soc_p, sha_p))
    nd_f, nd_s, nd_r = discriminate(f_pos, friend_v), discriminate(sh_pos, sharing_v), discriminate(r_pos, rec_u)
    total = rec_loss + ss_rate * (nd_f + nd_s + nd_r)
    return dict(rec_user=rec_u, rec_item=rec_i, sharing_view=sharing_v, friend_view=friend_v, rec_loss=rec_loss,
                social_prediction=soc_p, sharing_prediction=sha_p, rec_prediction=rec_p, f_pos=f_pos, sh_pos=sh_pos, r_pos=r_pos,
                nd_f=nd_f, nd_s=nd_s, nd_r=nd_r, total=total)


# ------------------------------------------------------------------------------------------ esrf.py
def esrf_generator(relation, selector, A, n_layers, segment, noise):
    """ESRF.Generator.forward (esrf.py:1127-1149) with the uniform draws of gumbel_softmax (esrf.py:1003-1008) given."""
    n = relation.shape[0]
    embs, x = [relation], relation
    for _ in range(n_layers):
        x = A @ x                                            # the raw product feeds the next layer
        embs.append(F.normalize(x, p=2, dim=1))
    emb = torch.stack(embs, 0).mean(0)
    end = min(segment + 100, n)
    feats = emb[segment:end] @ emb.T
    eps = 1e-10
    rows = []
    for r in range(feats.shape[0]):
        alpha = feats[r].unsqueeze(0) * selector
        g = -torch.log(-torch.log(noise[r] + eps) + eps)
        rows.append(F.softmax((torch.log(alpha + eps) + g) / 0.2, dim=-1).sum(0))
    alt = torch.zeros(n, n, dtype=relation.dtype)
    alt[segment:end] = torch.stack(rows)
    return alt


def esrf_discriminator(user_w, item_w, adj, alt, n_layers, is_social, K):
    """ESRF.Discriminator.forward (esrf.py:1168-1198)."""
    nu = user_w.shape[0]
    ego = torch.cat([user_w, item_w], 0)
    embs = [ego]
    for _ in range(n_layers):
        if is_social:
            ego = torch.cat([ego[:nu] + (alt @ ego[:nu]) / K, ego[nu:]], 0)
        else:
            ego = adj @ ego
        embs.append(F.normalize(ego, p=2, dim=1))
    total = torch.stack(embs, 0).sum(0)
    return total[:nu], total[nu:]


def esrf_losses(user_emb, item_emb, alt, u_idx, i_idx, j_idx, K, reg_u):
    """(pairwise, reg, adversarial, g_adv) of esrf.py:1231-1236 and 1296-1309."""
    ue, ve, ne = user_emb[u_idx], item_emb[i_idx], item_emb[j_idx]
    y_ui, y_uj = (ue * ve).sum(1), (ue * ne).sum(1)
    pair = -torch.sum(torch.log(torch.sigmoid(y_ui - y_uj) + 1e-10))
    reg = reg_u * (torch.norm(ue) + torch.norm(ve) + torch.norm(ne))
    y_vi = ((alt[u_idx] @ user_emb) / K * ve).sum(1)
    adv = -torch.sum(torch.log(torch.sigmoid(y_ui - y_vi) + 1e-10))
    g_adv = -torch.sum(torch.log(torch.sigmoid(y_vi - y_ui) + 1e-10))
    return pair, reg, adv, g_adv
