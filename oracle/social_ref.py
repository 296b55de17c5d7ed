"""Oracle: CPU restatement (torch fp64, dense adjacency) of the social-graph models' forward passes.  Test infrastructure only.

  mhcn_forward     univariate/mhcn.py:404-505 (self_gating, channel_attention, forward, hierarchical_self_supervision)
  diffnet_forward  univariate/diffnet.py:1124-1132
Pinned against tests/golden/{mhcn_model,diffnet_model}.npz, which the reference's own classes produced.
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
