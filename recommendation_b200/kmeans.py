"""GPU k-means for NCL's E-step (ncl.py:339-356).

The reference calls `faiss.Kmeans(d, k, gpu=False).train(x)` + `index.search(x, 1)` on the CPU inside the batch loop
(ncl.py:313,324: `self.e_step()` before the first and after every batch).  Here Lloyd's iterations run on the device:
the point-centroid distance block is a plain dense GEMM (cuBLAS through torch.matmul, fp32), the argmin a torch reduction,
and the centroid update the libgcf scatter-add kernel.  faiss defaults kept: 25 iterations, centroids initialised from a random
subset of the points, empty clusters re-seeded by splitting the largest one.  faiss's own RNG stream cannot be reproduced, so
only the algorithm (not the draw) is the reference behaviour; `init` lets a caller pin the start.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import functional as F_


def _assign(x: torch.Tensor, x_sq: torch.Tensor, c: torch.Tensor, block: int = 1 << 16) -> Tuple[torch.Tensor, torch.Tensor]:
    """Nearest centroid of every point: (index int64 [N], squared distance fp32 [N])."""
    c_sq = (c * c).sum(1)
    idx = torch.empty(x.shape[0], dtype=torch.int64, device=x.device)
    dist = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    for s in range(0, x.shape[0], block):
        e = min(x.shape[0], s + block)
        d2 = torch.addmm(c_sq.unsqueeze(0), x[s:e], c.T, beta=1.0, alpha=-2.0)   # |c|^2 - 2 x.c   (+ |x|^2 below)
        m, i = d2.min(dim=1)
        idx[s:e] = i
        dist[s:e] = (m + x_sq[s:e]).clamp_min_(0)
    return idx, dist


def kmeans(x: torch.Tensor, k: int, *, niter: int = 25, seed: int = 1234, init: Optional[torch.Tensor] = None):
    """Lloyd's algorithm.  Returns (centroids fp32 [k, d], assignment int64 [N], objective float = sum of squared distances)."""
    if not x.is_cuda:
        raise RuntimeError("kmeans needs a CUDA tensor: recommendation_b200 has no CPU path")
    x = x.detach().to(torch.float32).contiguous()
    n, d = x.shape
    if not 1 <= k <= n:
        raise ValueError("k must be in [1, number of points]")
    gen = torch.Generator(device=x.device)
    gen.manual_seed(seed)
    c = init.to(x.device, torch.float32).clone() if init is not None else x[torch.randperm(n, device=x.device, generator=gen)[:k]].clone()
    x_sq = (x * x).sum(1)
    ones = torch.ones(n, 1, dtype=torch.float32, device=x.device)
    for _ in range(niter):
        idx, _ = _assign(x, x_sq, c)
        sums = torch.zeros(k, d, dtype=torch.float32, device=x.device)
        F_.scatter_add_rows_(sums, idx, x)                                   # warp-aggregated red.add of the rows per cluster
        counts = torch.zeros(k, dtype=torch.float32, device=x.device).index_add_(0, idx, ones[:, 0])
        c = torch.where(counts.unsqueeze(1) > 0, sums / counts.clamp_min(1).unsqueeze(1), c)
        empty = (counts == 0).nonzero().flatten()
        if empty.numel():   # faiss: an empty cluster takes a slightly perturbed copy of a large cluster's centroid
            big = torch.argsort(counts, descending=True)[: empty.numel()]
            eps = 1.0 / 1024.0
            c[empty] = c[big] * (1 + eps)
            c[big] = c[big] * (1 - eps)
    idx, dist = _assign(x, x_sq, c)
    return c, idx, float(dist.sum())


def run_kmeans(x, k: int, *, seed: int = 1234) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """NCLModel.run_kmeans (ncl.py:347-356): x = embeddings ([n, d] tensor or numpy array); k is first capped at
    max(2, n // 39) as in the reference.  Returns (centroids on the device, assignment LongTensor [n], the k used)."""
    if not torch.is_tensor(x):
        x = torch.as_tensor(x)
    x = x.to(torch.device("cuda", torch.cuda.current_device()) if not x.is_cuda else x.device)
    k = min(k, max(2, x.shape[0] // 39))
    c, idx, _ = kmeans(x, k, seed=seed)
    return c, idx, k
