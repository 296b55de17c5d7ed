"""Synthetic power-law bipartite interaction graphs of the BASELINE.json config shapes.

Endpoints are drawn from Zipf-like weights p(k) ~ k^-alpha (alpha_user = 0.6, alpha_item = 0.8), pairs are
de-duplicated, truncated to exactly E and shuffled.  `numpy` flavour for tests / the CPU arm, `torch` flavour
(runs on the GPU) for the 100M-edge scaling graph.  Data generation is not part of any timed region.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

import numpy as np

# name -> (n_users, n_items, n_edges, d, n_layers)
CONFIGS = {
    "cfg1": (29_858, 40_981, 1_027_370, 64, 3),      # Gowalla-shaped, lightgcn.py
    "cfg2": (52_643, 91_599, 2_984_108, 64, 3),      # Amazon-book-shaped, ncl.py / ssl4rec.py
    "cfg3": (31_668, 38_048, 1_561_406, 128, 2),     # Yelp2018-shaped, directau.py / selfcf.py
    "cfg4": (250_000, 125_000, 5_000_000, 64, 2),    # social variants (mhcn.py / diffnet.py), +1M social edges
    "cfg5": (10_000_000, 5_000_000, 100_000_000, 64, 3),  # scaling sweep
    "tiny": (300, 500, 4_000, 64, 3),
    "small": (3_000, 4_000, 100_000, 64, 3),
}
ALPHA_USER, ALPHA_ITEM = 0.6, 0.8


@dataclass
class Interactions:
    n_users: int
    n_items: int
    users: np.ndarray  # int64 [E]
    items: np.ndarray  # int64 [E]

    @property
    def n_edges(self) -> int:
        return int(self.users.shape[0])


def _zipf_cdf(n: int, alpha: float) -> np.ndarray:
    w = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    cdf = np.cumsum(w)
    return cdf / cdf[-1]


def power_law_bipartite(n_users: int, n_items: int, n_edges: int, *, seed: int, alpha_user: float = ALPHA_USER,
                        alpha_item: float = ALPHA_ITEM) -> Interactions:
    """numpy generator (CPU).  Deterministic for a given seed."""
    if n_edges > n_users * n_items:
        raise ValueError("more edges than cells")
    rng = np.random.default_rng(seed)
    cu, ci = _zipf_cdf(n_users, alpha_user), _zipf_cdf(n_items, alpha_item)
    keys = np.empty(0, dtype=np.int64)
    while keys.shape[0] < n_edges:
        m = int((n_edges - keys.shape[0]) * 1.35) + 1024
        u = np.minimum(np.searchsorted(cu, rng.random(m)), n_users - 1).astype(np.int64)
        i = np.minimum(np.searchsorted(ci, rng.random(m)), n_items - 1).astype(np.int64)
        keys = np.unique(np.concatenate([keys, u * n_items + i]))
    keys = rng.permutation(keys)[:n_edges]
    keys = rng.permutation(keys)
    return Interactions(n_users, n_items, keys // n_items, keys % n_items)


def power_law_bipartite_torch(n_users: int, n_items: int, n_edges: int, *, seed: int, device,
                              alpha_user: float = ALPHA_USER, alpha_item: float = ALPHA_ITEM):
    """torch generator (runs where `device` is); returns (users, items) int64 tensors on `device`."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)

    def cdf(n, alpha):
        w = torch.arange(1, n + 1, dtype=torch.float64, device=device).pow_(-alpha)
        c = torch.cumsum(w, 0)
        return c / c[-1]

    cu, ci = cdf(n_users, alpha_user), cdf(n_items, alpha_item)
    keys = torch.empty(0, dtype=torch.int64, device=device)
    while keys.numel() < n_edges:
        m = int((n_edges - keys.numel()) * 1.35) + 1024
        u = torch.searchsorted(cu, torch.rand(m, dtype=torch.float64, device=device, generator=gen)).clamp_(max=n_users - 1)
        i = torch.searchsorted(ci, torch.rand(m, dtype=torch.float64, device=device, generator=gen)).clamp_(max=n_items - 1)
        keys = torch.unique(torch.cat([keys, u * n_items + i]))
        del u, i
    perm = torch.randperm(keys.numel(), device=device, generator=gen)[:n_edges]
    keys = keys[perm]
    return keys // n_items, keys % n_items


def config_graph(name: str, *, seed: int = None) -> Tuple[Interactions, int, int]:
    """(interactions, d, n_layers) for a named config; seed defaults to 1000 + cfg number."""
    n_users, n_items, n_edges, d, k = CONFIGS[name]
    if seed is None:
        seed = 1000 + (int(name[3:]) if name.startswith("cfg") else 0)
    return power_law_bipartite(n_users, n_items, n_edges, seed=seed), d, k
