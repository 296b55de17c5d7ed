"""Batched evaluation: top-N recommendation for all test users and the reference's ranking metrics.

Reference (ncl.py:133-178,253-277 = directau.py / selfcf.py / mhcn.py copies; lightgcn.py:48-74):
    test()               per user: predict (one matmul), rated items := -1e8, torch.topk(max_N), list of (item, score)
    ranking_evaluation() per N: Hit Ratio, Precision, Recall, NDCG over dicts of string ids
Here the score blocks are dense GEMMs (cuBLAS via torch.matmul, as in the reference), the masking / exact top-N selection /
hit counting run in libgcf kernels for a whole block of users at once, and the four measures are reduced on the device.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .graph import CSRGraph

MASK_VALUE = -1e8  # ncl.py:259


def positives_csr(users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(row_ptr, col_idx) int32: per-user sorted, de-duplicated item lists (training items to mask / test items to hit)."""
    g = CSRGraph.from_coo(users.to(torch.int64), items.to(torch.int64), None, n_users, n_items, norm="none")
    return g.row_ptr, g.col_idx


def recommend_topn(user_emb: torch.Tensor, item_emb: torch.Tensor, query_users: torch.Tensor, max_n: int, *,
                   train_pos: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                   block_bytes: int = 1 << 30) -> Tuple[torch.Tensor, torch.Tensor]:
    """Top-`max_n` items (ids int64 [Q, max_n], scores fp32 [Q, max_n]) for every user in `query_users`, training items
    excluded.  Score blocks of at most `block_bytes` are materialised, selected from and discarded."""
    lib = _lib.load()
    if not (user_emb.is_cuda and item_emb.is_cuda):
        raise RuntimeError("recommend_topn needs CUDA tensors: recommendation_b200 has no CPU path")
    dev = user_emb.device
    query_users = query_users.to(dev, torch.int64).contiguous()
    q, n_items = query_users.numel(), item_emb.shape[0]
    if not 1 <= max_n <= min(128, n_items):
        raise ValueError("max_n must be in [1, min(128, n_items)]")
    out_idx = torch.empty(q, max_n, dtype=torch.int64, device=dev)
    out_val = torch.empty(q, max_n, dtype=torch.float32, device=dev)
    rp, ci = train_pos if train_pos is not None else (None, None)
    rows_per_block = max(1, min(q, block_bytes // (4 * n_items)))
    item_t = item_emb.detach().to(torch.float32)
    for s in range(0, q, rows_per_block):
        e = min(q, s + rows_per_block)
        scores = torch.matmul(user_emb.detach()[query_users[s:e]].to(torch.float32), item_t.T)  # predict(), batched
        _lib.check(lib.gcf_masked_topn(_lib.ptr(scores), scores.stride(0), e - s, n_items, _lib.ptr(query_users[s:e]),
                                       _lib.ptr(rp), _lib.ptr(ci), MASK_VALUE, max_n, _lib.ptr(out_idx[s:e]),
                                       _lib.ptr(out_val[s:e]), _lib.current_stream()), "gcf_masked_topn")
    return out_idx, out_val


def hits_and_dcg(topn_idx: torch.Tensor, query_users: torch.Tensor, test_pos: Tuple[torch.Tensor, torch.Tensor],
                 cuts: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """(hits int32 [Q, len(cuts)], dcg fp32 [Q, len(cuts)]) at ascending cut-offs (gcf_ranking_hits)."""
    lib = _lib.load()
    dev = topn_idx.device
    q, n_top = topn_idx.shape
    cuts = [int(c) for c in cuts]
    if cuts != sorted(cuts) or cuts[-1] > n_top:
        raise ValueError("cut-offs must be ascending and not exceed the length of the recommendation lists")
    cut_t = torch.tensor(cuts, dtype=torch.int32, device=dev)
    hits = torch.empty(q, len(cuts), dtype=torch.int32, device=dev)
    dcg = torch.empty(q, len(cuts), dtype=torch.float32, device=dev)
    rp, ci = test_pos
    users = query_users.to(dev, torch.int64).contiguous()
    _lib.check(lib.gcf_ranking_hits(_lib.ptr(topn_idx.contiguous()), q, n_top, _lib.ptr(users), _lib.ptr(rp), _lib.ptr(ci),
                                    _lib.ptr(cut_t), len(cuts), _lib.ptr(hits), _lib.ptr(dcg), _lib.current_stream()),
               "gcf_ranking_hits")
    return hits, dcg


def ranking_measures(topn_idx: torch.Tensor, query_users: torch.Tensor, test_pos: Tuple[torch.Tensor, torch.Tensor],
                     top_ns: Sequence[int], *, n_test: Optional[torch.Tensor] = None) -> Dict[int, Dict[str, float]]:
    """{N: {'Hit Ratio', 'Precision', 'Recall', 'NDCG'}} with the definitions (and 5-decimal rounding) of Metric
    (ncl.py:133-162).  Every query user must have at least one test item (the reference iterates over test_set).

    `n_test` [Q] overrides len(origin[u]) -- the denominators of Hit Ratio / Recall and the length of the ideal list of
    NDCG.  The reference counts EVERY test item of the user there, also items that never occur in the training data and
    therefore have no column to be hit in (ncl.py:144,153,160); `test_pos` can only hold the mapped ones."""
    dev = topn_idx.device
    q, n_top = topn_idx.shape
    cuts = sorted(int(n) for n in top_ns)
    hits, dcg = hits_and_dcg(topn_idx, query_users, test_pos, cuts)
    rp, ci = test_pos
    users = query_users.to(dev, torch.int64).contiguous()
    if n_test is None:
        n_test = rp[users + 1] - rp[users]                        # len(origin[u]) when every test item has a column
    n_test = n_test.to(dev, torch.float64)
    idcg_table = torch.tensor([0.0] + list(np.cumsum([1.0 / math.log2(i + 2) for i in range(cuts[-1])])), dtype=torch.float64, device=dev)
    out: Dict[int, Dict[str, float]] = {}
    h64 = hits.to(torch.float64)
    for c, n in enumerate(cuts):
        idcg = idcg_table[torch.clamp(n_test, max=n).to(torch.int64)]
        ndcg = torch.where(idcg > 0, dcg[:, c].to(torch.float64) / idcg.clamp_min(1e-300), torch.zeros_like(idcg))
        out[n] = {"Hit Ratio": round(float(h64[:, c].sum() / n_test.sum()), 5),
                  "Precision": round(float(h64[:, c].sum() / (q * n)), 5),
                  "Recall": round(float((h64[:, c] / n_test).mean()), 5),
                  "NDCG": round(float(ndcg.sum() / q), 5)}
    return out


def format_measures(measures: Dict[int, Dict[str, float]]) -> List[str]:
    """The list-of-strings format ranking_evaluation returns (ncl.py:165-178)."""
    lines: List[str] = []
    for n in sorted(measures):
        lines.append(f"Top {n}\n")
        lines += [f"{k}:{measures[n][k]}\n" for k in ("Hit Ratio", "Precision", "Recall", "NDCG")]
    return lines


def evaluate_model(user_emb: torch.Tensor, item_emb: torch.Tensor, data, top_ns: Sequence[int]) -> List[str]:
    """test() + ranking_evaluation() for a reference `Interaction` object (`data.test_set`, `data.training_data`,
    id maps) in one call; returns the reference's result strings."""
    dev = user_emb.device
    tr_u = torch.tensor([data.user[r[0]] for r in data.training_data], dtype=torch.int64, device=dev)
    tr_i = torch.tensor([data.item[r[1]] for r in data.training_data], dtype=torch.int64, device=dev)
    # hits can only happen on items that have a column; the denominators count every test item (see ranking_measures)
    te = [(data.user[u], data.item[i]) for u, its in data.test_set.items() for i in its if i in data.item]
    te_u = torch.tensor([p[0] for p in te], dtype=torch.int64, device=dev)
    te_i = torch.tensor([p[1] for p in te], dtype=torch.int64, device=dev)
    train_pos = positives_csr(tr_u, tr_i, data.user_num, data.item_num)
    test_pos = positives_csr(te_u, te_i, data.user_num, data.item_num)
    query = torch.tensor(sorted(data.user[u] for u in data.test_set), dtype=torch.int64, device=dev)   # test(): every user of test_set
    raw_count = torch.zeros(data.user_num, dtype=torch.int64)
    for u, its in data.test_set.items():
        raw_count[data.user[u]] = len(its)
    idx, _ = recommend_topn(user_emb, item_emb, query, max(top_ns), train_pos=train_pos)
    return format_measures(ranking_measures(idx, query, test_pos, top_ns, n_test=raw_count.to(dev)[query]))
