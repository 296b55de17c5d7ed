"""Oracle: numpy restatement of the reference's evaluation (test(): mask + top-N, ncl.py:253-266; Metric, ncl.py:133-162).
Test infrastructure only."""
from __future__ import annotations

import math

import numpy as np


def masked_topn(scores, train_items_per_user, n_top, mask_value=-1e8):
    """scores [Q, I]; train_items_per_user: list of index arrays.  Ties broken by the lower item id."""
    s = np.array(scores, dtype=np.float32, copy=True)
    for q, its in enumerate(train_items_per_user):
        s[q, its] = mask_value
    order = np.lexsort((np.arange(s.shape[1])[None, :].repeat(s.shape[0], 0), -s.astype(np.float64)), axis=1)[:, :n_top]
    return order, np.take_along_axis(s, order, 1)


def measures(lists, test_items_per_user, top_ns):
    """{N: (hit_ratio, precision, recall, ndcg)} rounded to 5 decimals like Metric.*"""
    out = {}
    for n in top_ns:
        hits, dcgs = [], []
        for l, t in zip(lists, test_items_per_user):
            tset = set(int(x) for x in t)
            h = [int(x) in tset for x in l[:n]]
            hits.append(sum(h))
            dcg = sum(1.0 / math.log2(i + 2) for i, ok in enumerate(h) if ok)
            idcg = sum(1.0 / math.log2(i + 2) for i in range(min(len(tset), n)))
            dcgs.append(dcg / idcg if idcg else 0.0)
        total = sum(len(t) for t in test_items_per_user)
        out[n] = (round(sum(hits) / total, 5), round(sum(hits) / (len(lists) * n), 5),
                  round(float(np.mean([h / len(t) for h, t in zip(hits, test_items_per_user)])), 5), round(sum(dcgs) / len(lists), 5))
    return out


def lightgcn_evaluate(scores, train_pos, test_users, test_items, k_list):
    """lightgcn.py:48-74 restated: scores [U, I]; train_pos {user: set}; test pairs as arrays.  Returns {k: {HR, P, R, NDCG}}."""
    out = {k: {"HR": 0.0, "P": 0.0, "R": 0.0, "NDCG": 0.0} for k in k_list}
    users = np.unique(test_users)
    for u in users:
        s = np.array(scores[u], dtype=np.float64)
        known = list(train_pos.get(int(u), ()))
        if known:
            s[known] = -np.inf
        top = np.lexsort((np.arange(s.shape[0]), -s))[:max(k_list)]
        tset = set(int(x) for x in test_items[test_users == u])
        for k in k_list:
            hits = sum(1 for it in top[:k] if int(it) in tset)
            out[k]["HR"] += hits > 0
            out[k]["P"] += hits / k
            out[k]["R"] += hits / len(tset) if tset else 0
            out[k]["NDCG"] += sum(1 / np.log2(i + 2) for i in range(k) if int(top[i]) in tset)
    for k in k_list:
        for m in out[k]:
            out[k][m] /= len(users)
    return out
