"""Drop-in optimisers on the fused libgcf update kernels.

    torch.optim.Adam(model.parameters(), lr)                 ncl.py:305,329  selfcf.py:542  directau.py:212  lightgcn.py:80
    torch.optim.SGD(model.parameters(), lr, momentum=0.9)    selfcf.py:544   directau.py:214

`Adam` / `SGD` below take the same constructor arguments and keep the same `state_dict()` layout (`step`, `exp_avg`,
`exp_avg_sq` / `momentum_buffer`), so a checkpoint written by the reference's optimiser loads into them and vice versa.
Each parameter is updated by ONE kernel launch that reads and writes param / grad / state once.  Parameters that are not
contiguous float32 CUDA tensors are refused -- there is no eager fallback.
"""
from __future__ import annotations

import torch

from . import functional as F_


def _check(p: torch.Tensor) -> None:
    if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous()):
        raise TypeError("recommendation_b200.optim: parameters must be contiguous float32 CUDA tensors")
    if p.grad.is_sparse:
        raise RuntimeError("recommendation_b200.optim: sparse gradients are not supported (use F_.adam_rows_step_)")


class SGD(torch.optim.Optimizer):
    """torch.optim.SGD numerics (single-tensor update order), one fused pass per parameter (gcf_sgd_momentum_step)."""

    def __init__(self, params, lr: float = 1e-3, momentum: float = 0.0, dampening: float = 0.0, weight_decay: float = 0.0,
                 nesterov: bool = False):
        if lr < 0.0 or momentum < 0.0 or weight_decay < 0.0:
            raise ValueError("SGD: lr, momentum and weight_decay must be non-negative")
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError("Nesterov momentum requires a momentum and zero dampening")
        super().__init__(params, dict(lr=lr, momentum=momentum, dampening=dampening, weight_decay=weight_decay, nesterov=nesterov))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p)
                state = self.state[p]
                buf, first = None, False
                if group["momentum"] != 0.0:
                    buf = state.get("momentum_buffer")
                    if buf is None:
                        buf = state["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
                        first = True
                F_.sgd_step_(p, p.grad.contiguous(), buf, lr=group["lr"], momentum=group["momentum"], dampening=group["dampening"],
                             weight_decay=group["weight_decay"], nesterov=group["nesterov"], first_step=first)
        return loss


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam (decoupled=False) / AdamW (decoupled=True) numerics, one fused pass per parameter (gcf_adam_step)."""

    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 decoupled: bool = False):
        if lr < 0.0 or eps < 0.0 or not (0.0 <= betas[0] < 1.0) or not (0.0 <= betas[1] < 1.0):
            raise ValueError("Adam: invalid hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, decoupled=decoupled))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                _check(p)
                state = self.state[p]
                if not state:
                    state["step"] = torch.zeros((), dtype=torch.float32)       # host scalar, as torch keeps it (capturable=False)
                    state["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    state["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                state["step"] += 1
                F_.adam_step_(p, p.grad.contiguous(), state["exp_avg"], state["exp_avg_sq"], int(state["step"].item()), lr=group["lr"],
                              betas=group["betas"], eps=group["eps"], weight_decay=group["weight_decay"], decoupled=group["decoupled"])
        return loss
