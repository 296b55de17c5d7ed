"""GPU k-means for NCL's E-step (ncl.py:339-356).

The reference calls `faiss.Kmeans(d, k, gpu=False).train(x)` + `index.search(x, 1)` on the CPU inside the batch loop
(ncl.py:313,324: `self.e_step()` before the first and after every batch).  Here the whole Lloyd loop is ONE C-ABI call,
gcf_kmeans_lloyd (csrc/kmeans.cu): point-centroid products on the tcgen05 tensor cores with a running arg-min epilogue,
deterministic sorted centroid sums, empty clusters re-seeded on the device -- no host synchronisation per iteration.
faiss defaults kept: 25 iterations, centroids initialised from a random subset of the points, an empty cluster takes a
slightly perturbed copy of a large cluster's centroid.  faiss's own RNG stream cannot be reproduced, so only the algorithm
(not the draw) is the reference behaviour; `init` lets a caller pin the start.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _lib


def kmeans(x: torch.Tensor, k: int, *, niter: int = 25, seed: int = 1234, init: Optional[torch.Tensor] = None):
    """Lloyd's algorithm.  Returns (centroids fp32 [k, d], assignment int64 [N], objective float = sum of squared distances)."""
    if not x.is_cuda:
        raise RuntimeError("kmeans needs a CUDA tensor: recommendation_b200 has no CPU path")
    lib = _lib.load()
    x = x.detach().to(torch.float32)
    if x.dim() != 2:
        raise ValueError("kmeans: x must be [n, d]")
    if x.stride(1) != 1 or x.stride(0) % 4 != 0 or x.data_ptr() % 16 != 0:
        x = x.contiguous()
    n, d = x.shape
    if not 1 <= k <= n:
        raise ValueError("k must be in [1, number of points]")
    if init is not None:
        c = init.to(x.device, torch.float32).contiguous().clone()
        if c.shape != (k, d):
            raise ValueError("kmeans: init must be [k, d]")
    else:
        gen = torch.Generator(device=x.device)
        gen.manual_seed(seed)
        c = x[torch.randperm(n, device=x.device, generator=gen)[:k]].contiguous().clone()
    idx = torch.empty(n, dtype=torch.int64, device=x.device)
    dist = torch.empty(n, dtype=torch.float32, device=x.device)
    ws_bytes = lib.gcf_kmeans_workspace_bytes(n, k, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    _lib.check(lib.gcf_kmeans_lloyd(_lib.ptr(x), x.stride(0), n, d, k, int(niter), _lib.ptr(c), _lib.ptr(idx), _lib.ptr(dist),
                                    _lib.ptr(ws), ws_bytes, _lib.current_stream()), "gcf_kmeans_lloyd")
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
