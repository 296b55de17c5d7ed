"""Device-side replacement of the reference's Python batch iterators.

    next_batch_pairwise(data, batch_size, n_negs=1)     ncl.py:91-114 (= directau.py:14-32, selfcf.py:188-211,
                                                         ssl4rec.py:33-50 up to the trial cap)
yields (u_idx, i_idx, j_idx) per batch: shuffled training pairs and, for each, an item drawn uniformly from all
items that is NOT one of the user's training items (rejection, at most `max_trials` redraws -- ncl.py caps at 100).
The reference builds Python lists one `random.choice` at a time (3.4 k samples/s); here the
shuffle is one randperm, the negatives one Philox kernel launch per batch, and the index tensors stay on the GPU
(every reference call site only uses them to index embedding tables).
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import numpy as np
import torch

from . import functional as F_
from .graph import CSRGraph


class PairwiseSampler:
    """Training pairs + per-user sorted positives CSR on the device."""

    def __init__(self, users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int, *, seed: int = 0):
        if not users.is_cuda:
            raise RuntimeError("PairwiseSampler needs CUDA tensors: recommendation_b200 has no CPU path")
        self.users = users.to(torch.int64).contiguous()
        self.items = items.to(torch.int64).contiguous()
        self.n_users, self.n_items, self.seed = n_users, n_items, seed
        pos = CSRGraph.from_coo(self.users, self.items, None, n_users, n_items, norm="none")  # sorted, de-duplicated
        self.pos_row_ptr, self.pos_col_idx = pos.row_ptr, pos.col_idx
        self.epoch = 0
        self._gen = torch.Generator(device=self.users.device)
        self._gen.manual_seed(seed)

    @classmethod
    def from_data(cls, data, *, seed: int = 0, device: Optional[torch.device] = None) -> "PairwiseSampler":
        """`data`: the reference's Interaction object (training_data rows [user, item, rating], id maps .user / .item)."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        u = np.fromiter((data.user[r[0]] for r in data.training_data), dtype=np.int64, count=len(data.training_data))
        i = np.fromiter((data.item[r[1]] for r in data.training_data), dtype=np.int64, count=len(data.training_data))
        return cls(torch.from_numpy(u).to(dev), torch.from_numpy(i).to(dev), data.user_num, data.item_num, seed=seed)

    def batches(self, batch_size: int, n_negs: int = 1, *, max_trials: int = 100,
                shuffle: bool = True) -> Iterator[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]]:
        n = self.users.numel()
        perm = torch.randperm(n, device=self.users.device, generator=self._gen) if shuffle else None
        self.epoch += 1
        for b, ptr in enumerate(range(0, n, batch_size)):
            sel = perm[ptr:ptr + batch_size] if perm is not None else slice(ptr, ptr + batch_size)
            u, i = self.users[sel], self.items[sel]
            j = F_.sample_negatives(u.numel(), self.n_items, seed=self.seed, offset=(self.epoch << 32) | b, n_negs=n_negs,
                                    users=u, positives=(self.pos_row_ptr, self.pos_col_idx), max_trials=max_trials)
            yield u, i, j


def next_batch_pairwise(data, batch_size: int, n_negs: int = 1):
    """Reference signature.  The sampler state (device pairs, positives CSR, epoch counter) is cached on `data`."""
    s = getattr(data, "_gcf_sampler", None)
    if s is None:
        # ingest.DeviceInteraction already holds the dense index tensors; the reference's Interaction needs its dictionaries walked
        s = data.sampler() if hasattr(data, "sampler") else PairwiseSampler.from_data(data)
        data._gcf_sampler = s
    yield from s.batches(batch_size, n_negs)
