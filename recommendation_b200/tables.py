"""User and item embedding tables in ONE [U+I, d] allocation.

Every reference encoder starts with `torch.cat([user_emb, item_emb], 0)` (ncl.py:416, selfcf.py:476, lightgcn.py:22) and
autograd ends with the matching split copy.  Here the two `nn.Parameter`s are views of one buffer, so the joint table
the propagation kernels need exists without a copy, and its gradient is handed back as two views.
"""
from __future__ import annotations

import torch
import torch.nn as nn


def join_parameters(user_p: nn.Parameter, item_p: nn.Parameter) -> torch.Tensor:
    """Re-seat both parameters on one [U+I, d] allocation (values preserved); returns the joint table."""
    u, d = user_p.shape
    i = item_p.shape[0]
    table = torch.empty(u + i, d, dtype=user_p.dtype, device=user_p.device)
    with torch.no_grad():
        table[:u].copy_(user_p)
        table[u:].copy_(item_p)
        user_p.data = table[:u]
        item_p.data = table[u:]
    return table


class JoinTables(torch.autograd.Function):
    """cat([user_w, item_w]) without the copy when both weights are adjacent slices of one allocation."""

    @staticmethod
    def forward(ctx, user_w: torch.Tensor, item_w: torch.Tensor):
        u, d = user_w.shape
        i = item_w.shape[0]
        ctx.split = (u, i)
        adjacent = (user_w.is_contiguous() and item_w.is_contiguous()
                    and user_w.untyped_storage().data_ptr() == item_w.untyped_storage().data_ptr()
                    and user_w.data_ptr() + u * d * 4 == item_w.data_ptr())
        if adjacent:
            return torch.as_strided(user_w.detach(), (u + i, d), (d, 1))
        return torch.cat([user_w, item_w], dim=0)

    @staticmethod
    def backward(ctx, g):
        u, i = ctx.split
        return g[:u], g[u:]


class JointEmbeddingDict(nn.Module):
    """Mixin for the `embedding_dict = ParameterDict({'user_emb', 'item_emb'})` encoders (ncl.py:407-413,
    selfcf.py:467-473): keeps both parameters on one allocation across .to() / load_state_dict()."""

    def _join(self) -> None:
        self._table = join_parameters(self.embedding_dict["user_emb"], self.embedding_dict["item_emb"])

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._join()
        return out

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._join()
        return out

    def joint_table(self) -> torch.Tensor:
        return JoinTables.apply(self.embedding_dict["user_emb"], self.embedding_dict["item_emb"])


def xavier_uniform_table(n_users: int, n_items: int, d: int, *, seed: int, device, cols=None, chunk_rows: int = 1 << 20) -> torch.Tensor:
    """Joint [U+I, d] table with the reference's initialisation (nn.init.xavier_uniform_ on the [U, d] and the [I, d]
    table separately: lightgcn.py:14-17, ncl.py:409-412), as a pure function of (seed, row, column).

    Rows are generated full-width in fixed chunks of `chunk_rows`, chunk c from its own generator seeded with (seed, c), and
    `cols` = (lo, hi) keeps only that column slice -- so every rank of a feature- or row-sharded run holds exactly the
    values the single-GPU run holds for the same entries, whatever the world size (the per-rank seeds of the
    sharded trainers repeated rows across shards)."""
    lo, hi = (0, d) if cols is None else cols
    n = n_users + n_items
    out = torch.empty(n, hi - lo, dtype=torch.float32, device=device)
    bound_u = (6.0 / (n_users + d)) ** 0.5
    bound_i = (6.0 / (n_items + d)) ** 0.5
    gen = torch.Generator(device=device)
    for c, r0 in enumerate(range(0, n, chunk_rows)):
        r1 = min(n, r0 + chunk_rows)
        gen.manual_seed((int(seed) * 1_000_003 + c) & 0x7FFFFFFFFFFFFFFF)
        t = torch.rand(r1 - r0, d, device=device, generator=gen)[:, lo:hi].mul_(2.0).sub_(1.0)
        cut = min(max(n_users - r0, 0), r1 - r0)     # rows of this chunk that are users
        t[:cut] *= bound_u
        t[cut:] *= bound_i
        out[r0:r1] = t
    return out
