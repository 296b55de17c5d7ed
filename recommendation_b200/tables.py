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
