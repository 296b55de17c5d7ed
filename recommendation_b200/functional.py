"""torch.autograd bridges from the reference's tensor-level operations to the libgcf kernels.

Each function names the reference op chain it replaces.  Tensors are fp32, contiguous, on CUDA;
anything else raises -- there is no eager fallback.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from .graph import CSRGraph

_DETERMINISTIC = os.environ.get("GCF_DETERMINISTIC", "0") == "1"


def set_deterministic(flag: bool) -> None:
    """Route gather backward through the sorted (order-reproducible) scatter-add."""
    global _DETERMINISTIC
    _DETERMINISTIC = bool(flag)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (got {t.device}); recommendation_b200 has no CPU path")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32 (got {t.dtype})")
    return t if t.is_contiguous() else t.contiguous()


def _rows_view(t: torch.Tensor, name: str) -> Tuple[torch.Tensor, int]:
    """Accept [n, d] tensors whose rows are contiguous (stride(1) == 1); returns (tensor, leading dim)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (got {t.device})")
    if t.dtype != torch.float32 or t.dim() != 2:
        raise TypeError(f"{name} must be a 2-D float32 tensor")
    if t.stride(1) != 1 or t.stride(0) < t.shape[1] or t.stride(0) % 4 != 0 or t.data_ptr() % 16 != 0:
        t = t.contiguous()
    return t, t.stride(0)


def _idx(t, device, name: str) -> torch.Tensor:
    """Reference call sites pass Python lists or LongTensors (ncl.py:314-316); normalise to int64 CUDA."""
    if not torch.is_tensor(t):
        t = torch.as_tensor(t, dtype=torch.int64)
    if t.dtype != torch.int64:
        t = t.to(torch.int64)
    if t.device != device:
        t = t.to(device, non_blocking=True)
    return t.contiguous()


# =========================================================================================
# SpMM / propagation
# =========================================================================================
def spmm_raw(graph: CSRGraph, x: torch.Tensor, *, y: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
             epilogue: int = _lib.EPILOGUE_NONE, alpha: float = 1.0, post: float = 1.0,
             addends: Sequence[torch.Tensor] = (), betas: Sequence[float] = (), variant: int = 0) -> None:
    """Direct (non-autograd) call of gcf_spmm_csr_f32; see include/gcf.h for the epilogue algebra."""
    lib = _lib.load()
    d = x.shape[1]
    ws, ws_bytes = graph.workspace(d)
    _lib.check(lib.gcf_spmm_csr_f32(graph.struct_ref(), d, _lib.ptr(x), x.stride(0),
                                    _lib.ptr(y), y.stride(0) if y is not None else 0,
                                    _lib.ptr(out), out.stride(0) if out is not None else 0,
                                    epilogue, alpha, post, len(addends), _lib.ptr_array(list(addends)),
                                    _lib.float_array(list(betas)), _lib.ptr(ws), ws_bytes, variant,
                                    _lib.current_stream()), "gcf_spmm_csr_f32")


class _SpMM(torch.autograd.Function):
    """Y = A @ X   (torch.sparse.mm, ncl.py:419 / selfcf.py:479 / mhcn.py:440-456)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, graph: CSRGraph):
        x = _f32c(x, "x")
        if x.shape[0] != graph.n_cols:
            raise ValueError(f"spmm: X has {x.shape[0]} rows, operator has {graph.n_cols} columns")
        y = torch.empty(graph.n_rows, x.shape[1], dtype=torch.float32, device=x.device)
        spmm_raw(graph, x, y=y)
        ctx.graph = graph
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = _f32c(gy, "grad")
        gt = ctx.graph.transpose()
        gx = torch.empty(gt.n_rows, gy.shape[1], dtype=torch.float32, device=gy.device)
        spmm_raw(gt, gy, y=gx)
        return gx, None


def spmm(graph: CSRGraph, x: torch.Tensor) -> torch.Tensor:
    return _SpMM.apply(x, graph)


class _Propagate(torch.autograd.Function):
    """K-layer LightGCN propagation with fused layer combination (gcf_propagate_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x0: torch.Tensor, graph: CSRGraph, n_layers: int, scale: float, want_layers: bool):
        lib = _lib.load()
        x0 = _f32c(x0, "x0")
        if graph.n_rows != graph.n_cols or x0.shape[0] != graph.n_rows:
            raise ValueError("propagate: operator must be square and match x0's row count")
        n, d = x0.shape
        layers: List[Optional[torch.Tensor]] = [torch.empty_like(x0) for _ in range(n_layers - 1)]
        layers.append(torch.empty_like(x0) if want_layers else None)
        final = torch.empty_like(x0)
        ws, ws_bytes = graph.workspace(d)
        _lib.check(lib.gcf_propagate_fwd(graph.struct_ref(), d, n_layers, _lib.ptr(x0), _lib.ptr_array(layers),
                                         _lib.ptr(final), scale, _lib.ptr(ws), ws_bytes, _lib.current_stream()),
                   "gcf_propagate_fwd")
        ctx.graph, ctx.n_layers, ctx.scale, ctx.want_layers = graph, n_layers, scale, want_layers
        ctx.set_materialize_grads(False)  # unused layer outputs arrive as None, not as zero tensors
        if want_layers:
            return (final, *layers)
        return (final,)

    @staticmethod
    def backward(ctx, g_final, *g_layers):
        lib = _lib.load()
        graph_t = ctx.graph.transpose()
        k = ctx.n_layers
        extra: List[Optional[torch.Tensor]] = [None] * (k + 1)
        ref = g_final
        for i, g in enumerate(g_layers):
            if g is not None:
                extra[i + 1] = _f32c(g, "layer grad")
                ref = ref if ref is not None else g
        if ref is None:
            return None, None, None, None, None
        if g_final is not None:
            g_final = _f32c(g_final, "grad")
        elif extra[k] is None:
            g_final = torch.zeros_like(ref)  # rare: gradient only reaches an inner layer output
        n, d = ref.shape
        ping = torch.empty(n, d, dtype=torch.float32, device=ref.device) if (k > 1 or extra[k] is not None) else None
        pong = torch.empty(n, d, dtype=torch.float32, device=ref.device) if k > 1 else None
        g_x0 = torch.empty(n, d, dtype=torch.float32, device=ref.device)
        ws, ws_bytes = graph_t.workspace(d)
        _lib.check(lib.gcf_propagate_bwd(graph_t.struct_ref(), d, k, _lib.ptr(g_final), _lib.ptr_array(extra),
                                         ctx.scale, _lib.ptr(ping), _lib.ptr(pong), _lib.ptr(g_x0), _lib.ptr(ws),
                                         ws_bytes, _lib.current_stream()), "gcf_propagate_bwd")
        # E(0) = x0 receives scale * g_final directly: already folded into G(0) by the kernel epilogue
        return g_x0, None, None, None, None


def propagate(graph: CSRGraph, x0: torch.Tensor, n_layers: int, *, mode: str = "mean",
              return_layers: bool = False):
    """final = mean|sum over [E0, A E0, ..., A^K E0].

    mode="mean": LGCNEncoder / LGCN_Encoder (ncl.py:415-422, selfcf.py:475-485, directau.py:286-293)
    mode="sum" : LightGCN.forward's `x += out` (lightgcn.py:21-27)
    return_layers=True also returns [E1..EK] (E0 is x0 itself), as in ncl.py:417-422.
    """
    if mode not in ("mean", "sum"):
        raise ValueError("mode must be 'mean' or 'sum'")
    if n_layers < 1:
        raise ValueError("n_layers must be >= 1")
    scale = 1.0 / (n_layers + 1) if mode == "mean" else 1.0
    outs = _Propagate.apply(x0, graph, int(n_layers), float(scale), bool(return_layers))
    if return_layers:
        return outs[0], list(outs[1:])
    return outs[0]


# =========================================================================================
# gather / scatter-add
# =========================================================================================
def scatter_add_rows_(table_grad: torch.Tensor, idx: torch.Tensor, src: torch.Tensor, *, deterministic: Optional[bool] = None) -> None:
    """table_grad[idx[t]] += src[t]   (index_put_(accumulate=True), the backward of x[idx])."""
    lib = _lib.load()
    det = _DETERMINISTIC if deterministic is None else deterministic
    src, lds = _rows_view(src, "src")
    n, d = src.shape
    mode = 1 if det else 0
    ws_bytes = lib.gcf_scatter_add_workspace_bytes(n, table_grad.shape[0], mode)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=src.device) if ws_bytes else None
    _lib.check(lib.gcf_scatter_add_rows(_lib.ptr(src), lds, d, _lib.ptr(idx), n, _lib.ptr(table_grad),
                                        table_grad.stride(0), table_grad.shape[0], mode, _lib.ptr(ws), ws_bytes,
                                        _lib.current_stream()), "gcf_scatter_add_rows")


class _GatherRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, table: torch.Tensor, idx: torch.Tensor):
        lib = _lib.load()
        table, ld = _rows_view(table, "table")
        n_rows, d = table.shape
        out = torch.empty(idx.numel(), d, dtype=torch.float32, device=table.device)
        _lib.check(lib.gcf_gather_rows(_lib.ptr(table), ld, n_rows, d, _lib.ptr(idx), idx.numel(), _lib.ptr(out), d,
                                       _lib.current_stream()), "gcf_gather_rows")
        ctx.save_for_backward(idx)
        ctx.shape = (n_rows, d)
        return out

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        g = _f32c(g, "grad")
        table_grad = torch.zeros(ctx.shape, dtype=torch.float32, device=g.device)
        scatter_add_rows_(table_grad, idx, g)
        return table_grad, None


def gather_rows(table: torch.Tensor, idx) -> torch.Tensor:
    """table[idx] with a warp-aggregated scatter-add backward (ncl.py:314-316, selfcf.py:504-511, ...)."""
    idx = _idx(idx, table.device, "idx")
    if os.environ.get("GCF_CHECK_INDEX") == "1" and idx.numel():  # costs a device sync; off by default
        if int(idx.min()) < 0 or int(idx.max()) >= table.shape[0]:
            raise IndexError("gather_rows: index out of range")
    return _GatherRows.apply(table, idx)


# =========================================================================================
# negative sampler
# =========================================================================================
def sample_negatives(n: int, n_items: int, *, seed: int, offset: int = 0, n_negs: int = 1,
                     users: Optional[torch.Tensor] = None, positives: Optional[Tuple[torch.Tensor, torch.Tensor]] = None,
                     max_trials: int = 100, device=None) -> torch.Tensor:
    """Philox4x32-10 negatives; with `positives=(row_ptr, col_idx)` (per-user sorted item CSR) candidates
    that are training positives of `users[t]` are rejected (ncl.py:91-114), otherwise uniform as
    torch.randint (lightgcn.py:91-94).  Returns int64 [n] (n_negs == 1) or [n, n_negs]."""
    lib = _lib.load()
    dev = device if device is not None else (users.device if users is not None else torch.device("cuda", torch.cuda.current_device()))
    out = torch.empty(n * n_negs, dtype=torch.int64, device=dev)
    rp = ci = None
    if positives is not None:
        rp, ci = positives
        if users is None:
            raise ValueError("rejection sampling needs the user of each triple")
        users = _idx(users, dev, "users")
    _lib.check(lib.gcf_sample_negatives(int(seed) & (2**64 - 1), int(offset) & (2**64 - 1), _lib.ptr(users), n, n_negs,
                                        n_items, _lib.ptr(rp), _lib.ptr(ci), max_trials, _lib.ptr(out),
                                        _lib.current_stream()), "gcf_sample_negatives")
    return out if n_negs == 1 else out.view(n, n_negs)


# =========================================================================================
# fused BPR
# =========================================================================================
class _BprFused(torch.autograd.Function):
    """Fused gather + BPR.  When a gradient is needed the backward is computed IN the forward launch
    (gcf_bpr_fwd_bwd: one pass over the triples, rows gathered once) with an upstream gradient of 1; backward()
    then only applies the actual upstream scalar, which is a no-op launch for `loss.backward()`."""

    @staticmethod
    def forward(ctx, user_emb, item_emb, u_idx, p_idx, n_idx, n_negs, variant, eps, reduction, reg_u, reg_p, reg_n, split=0):
        lib = _lib.load()
        joint = item_emb is None  # user_emb is the joint [U+I, d] table: users = rows [0, split), items = the rest
        if joint:
            table = _f32c(user_emb, "table")
            user_emb, item_emb = table[:split], table[split:]
            ldu = ldi = table.shape[1]
        else:
            user_emb, ldu = _rows_view(user_emb, "user_emb")
            item_emb, ldi = _rows_view(item_emb, "item_emb")
        d = user_emb.shape[1]
        if item_emb.shape[1] != d:
            raise ValueError("user/item embedding widths differ")
        n = u_idx.numel()
        dev = user_emb.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        ws_bytes = lib.gcf_bpr_workspace_bytes(n)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        need_grad = ctx.needs_input_grad[0] or (not joint and ctx.needs_input_grad[1])
        ctx.joint = joint
        ctx.consumed = False
        if need_grad:
            if joint:
                g_table = torch.zeros(table.shape, dtype=torch.float32, device=dev)
                g_user, g_item = g_table[:split], g_table[split:]
            else:
                g_user = torch.zeros(user_emb.shape, dtype=torch.float32, device=dev)
                g_item = torch.zeros(item_emb.shape, dtype=torch.float32, device=dev)
            _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(user_emb), ldu, _lib.ptr(item_emb), ldi, d, _lib.ptr(u_idx),
                                           _lib.ptr(p_idx), _lib.ptr(n_idx), n, n_negs, variant, eps, reduction,
                                           reg_u, reg_p, reg_n, 1.0, _lib.ptr(loss), None, _lib.ptr(g_user), d,
                                           _lib.ptr(g_item), d, _lib.ptr(ws), ws_bytes, _lib.current_stream()),
                       "gcf_bpr_fwd_bwd")
            ctx.grads = (g_table, None) if joint else (g_user, g_item)
        else:
            coef = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
            _lib.check(lib.gcf_bpr_fwd(_lib.ptr(user_emb), ldu, _lib.ptr(item_emb), ldi, d, _lib.ptr(u_idx), _lib.ptr(p_idx),
                                       _lib.ptr(n_idx), n, n_negs, variant, eps, reduction, reg_u, reg_p, reg_n,
                                       _lib.ptr(loss), _lib.ptr(coef), _lib.ptr(ws), ws_bytes, _lib.current_stream()),
                       "gcf_bpr_fwd")
            ctx.grads = None
        return loss

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        if ctx.grads is None:
            return (None,) * 13
        if ctx.consumed:
            raise RuntimeError("fused BPR gradients were already consumed (backward through this loss twice)")
        ctx.consumed = True
        g = g.contiguous().to(torch.float32)
        g_user, g_item = ctx.grads
        ctx.grads = None
        for t in (g_user, g_item):
            if t is not None:
                _lib.check(lib.gcf_scale_by_device_scalar(_lib.ptr(t), t.numel(), _lib.ptr(g), _lib.current_stream()),
                           "gcf_scale_by_device_scalar")
        return (g_user, g_item) + (None,) * 11


def _adjacent_views(a: torch.Tensor, b: torch.Tensor) -> Optional[torch.Tensor]:
    """If a = base[:U] and b = base[U:] of one contiguous [U+I, d] autograd tensor, return that base -- the
    (x[:U], x[U:]) pair LightGCN.forward returns (lightgcn.py:27).  Using the base directly keeps the gradient
    in one [U+I, d] buffer instead of two slice-backward zero-fills + copies."""
    base = a._base
    if base is None or b._base is not base or base.dim() != 2 or not base.is_contiguous():
        return None
    d = base.shape[1]
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != d or b.shape[1] != d or not (a.is_contiguous() and b.is_contiguous()):
        return None
    if a.data_ptr() != base.data_ptr() or b.data_ptr() != base.data_ptr() + a.shape[0] * d * 4:
        return None
    if a.shape[0] + b.shape[0] != base.shape[0]:
        return None
    return base


def bpr_loss_gather(user_emb: torch.Tensor, item_emb: torch.Tensor, u_idx, p_idx, n_idx, *,
                    variant: str = "softplus", eps: float = 1e-5, reduction: str = "mean",
                    reg_u: float = 0.0, reg_p: float = 0.0, reg_n: float = 0.0) -> torch.Tensor:
    """reduce_t l(<u,p> - mean_j <u,n_j>) + reg_u*sum|u|^2 + reg_p*sum|p|^2 + reg_n*sum|n|^2 with the three
    gathers fused in.  variant "log_eps_sigmoid" = ncl.py:116-120; "softplus" = lightgcn.py:108 / gcl.py:221."""
    dev = user_emb.device
    u_idx, p_idx, n_idx = _idx(u_idx, dev, "u_idx"), _idx(p_idx, dev, "p_idx"), _idx(n_idx, dev, "n_idx")
    n = u_idx.numel()
    if p_idx.numel() != n or n_idx.numel() % max(n, 1) != 0:
        raise ValueError("u_idx / p_idx / n_idx lengths are inconsistent")
    n_negs = n_idx.numel() // n if n else 1
    var = {"log_eps_sigmoid": _lib.BPR_LOG_EPS_SIGMOID, "softplus": _lib.BPR_SOFTPLUS}[variant]
    red = {"mean": _lib.REDUCE_MEAN, "sum": _lib.REDUCE_SUM}[reduction]
    base = _adjacent_views(user_emb, item_emb)
    if base is not None:  # one joint table: a single [U+I, d] gradient buffer is produced
        return _BprFused.apply(base, None, u_idx, p_idx, n_idx.reshape(-1), n_negs, var, float(eps), red,
                               float(reg_u), float(reg_p), float(reg_n), int(user_emb.shape[0]))
    return _BprFused.apply(user_emb, item_emb, u_idx, p_idx, n_idx.reshape(-1), n_negs, var, float(eps), red,
                           float(reg_u), float(reg_p), float(reg_n))


def bpr_loss_rows(user_rows: torch.Tensor, pos_rows: torch.Tensor, neg_rows: torch.Tensor, *,
                  variant: str = "log_eps_sigmoid", eps: float = 1e-5, reduction: str = "mean") -> torch.Tensor:
    """BPR on already-gathered [B, d] rows -- the reference's bpr_loss(user_emb, pos_item_emb, neg_item_emb)
    signature (ncl.py:116-120).  Runs the same fused kernel with identity indices."""
    b = user_rows.shape[0]
    items = torch.cat([pos_rows, neg_rows], dim=0)
    ar = torch.arange(b, dtype=torch.int64, device=user_rows.device)
    return bpr_loss_gather(user_rows, items, ar, ar, ar + b, variant=variant, eps=eps, reduction=reduction)


# =========================================================================================
# optimiser
# =========================================================================================
def adam_step_(param: torch.Tensor, grad: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int, *,
               lr: float, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, decoupled: bool = False) -> None:
    lib = _lib.load()
    for name, t in (("param", param), ("grad", grad), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"adam_step_: {name} must be a contiguous float32 CUDA tensor")
    _lib.check(lib.gcf_adam_step(_lib.ptr(param), _lib.ptr(grad), _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq), param.numel(),
                                 lr, betas[0], betas[1], eps, weight_decay, 1 if decoupled else 0, int(step),
                                 _lib.current_stream()), "gcf_adam_step")


def sgd_step_(param: torch.Tensor, grad: torch.Tensor, momentum_buf: Optional[torch.Tensor], *, lr: float, momentum: float = 0.0,
              dampening: float = 0.0, weight_decay: float = 0.0, nesterov: bool = False, first_step: bool = False) -> None:
    """One fused torch.optim.SGD update (selfcf.py:544, directau.py:214: momentum = 0.9).  `first_step` = the momentum
    buffer does not exist yet (torch initialises it with the gradient); it is written, not read, on that step."""
    lib = _lib.load()
    tensors = [("param", param), ("grad", grad)] + ([("momentum_buf", momentum_buf)] if momentum_buf is not None else [])
    for name, t in tensors:
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError(f"sgd_step_: {name} must be a contiguous float32 CUDA tensor")
    if momentum != 0.0 and momentum_buf is None:
        raise ValueError("sgd_step_: momentum needs a momentum buffer")
    _lib.check(lib.gcf_sgd_momentum_step(_lib.ptr(param), _lib.ptr(grad), _lib.ptr(momentum_buf), param.numel(), lr, momentum,
                                         dampening, weight_decay, 1 if nesterov else 0, 1 if first_step else 0,
                                         _lib.current_stream()), "gcf_sgd_momentum_step")


def adam_rows_step_(param: torch.Tensor, rows: torch.Tensor, grad_rows: torch.Tensor, exp_avg: torch.Tensor,
                    exp_avg_sq: torch.Tensor, step: int, *, lr: float, betas=(0.9, 0.999), eps: float = 1e-8,
                    weight_decay: float = 0.0, decoupled: bool = False) -> None:
    """Row-sparse Adam: rows[r] of param / moments updated with grad_rows[r]; `rows` must be distinct.
    torch.optim.SparseAdam semantics -- the moments of untouched rows do not decay."""
    lib = _lib.load()
    for name, t in (("param", param), ("grad_rows", grad_rows), ("exp_avg", exp_avg), ("exp_avg_sq", exp_avg_sq)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and t.is_contiguous()):
            raise TypeError(f"adam_rows_step_: {name} must be a contiguous 2-D float32 CUDA tensor")
    rows = _idx(rows, param.device, "rows")
    if grad_rows.shape[0] != rows.numel() or grad_rows.shape[1] != param.shape[1]:
        raise ValueError("adam_rows_step_: grad_rows must be [len(rows), d]")
    d = param.shape[1]
    _lib.check(lib.gcf_adam_rows_step(_lib.ptr(param), d, _lib.ptr(grad_rows), d, _lib.ptr(exp_avg), _lib.ptr(exp_avg_sq),
                                      _lib.ptr(rows), rows.numel(), d, lr, betas[0], betas[1], eps, weight_decay,
                                      1 if decoupled else 0, int(step), _lib.current_stream()), "gcf_adam_rows_step")


# =========================================================================================
# InfoNCE family (tcgen05 tensor cores)
# =========================================================================================
def infonce_stats_raw(q: torch.Tensor, k: torch.Tensor, tau: float, *, cos: bool = True, pos_idx: Optional[torch.Tensor] = None,
                      want_row: bool = True, want_col: bool = False, want_pos: bool = True):
    """(row_lse[M], col_lse[N] | None, pos[M]) of S = q^ k^T / tau without materialising S (gcf_infonce_fwd)."""
    lib = _lib.load()
    q, ldq = _rows_view(q, "q")
    k, ldk = _rows_view(k, "k")
    m, d = q.shape
    n = k.shape[0]
    if k.shape[1] != d:
        raise ValueError("q and k must have the same width")
    dev = q.device
    row = torch.empty(m, dtype=torch.float32, device=dev) if want_row else None
    col = torch.empty(n, dtype=torch.float32, device=dev) if want_col else None
    pos = torch.empty(m, dtype=torch.float32, device=dev) if want_pos else None
    ws_bytes = lib.gcf_infonce_workspace_bytes(m, n, d)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    if pos_idx is not None:
        pos_idx = _idx(pos_idx, dev, "pos_idx")
    _lib.check(lib.gcf_infonce_fwd(_lib.ptr(q), ldq, m, _lib.ptr(k), ldk, n, d, 1 if cos else 0, float(tau), _lib.ptr(pos_idx),
                                   _lib.ptr(row), _lib.ptr(col), _lib.ptr(pos), _lib.ptr(ws), ws_bytes, _lib.current_stream()),
               "gcf_infonce_fwd")
    return row, col, pos


class _InfoNCEStats(torch.autograd.Function):
    """(row_lse, col_lse, pos) of S = q^ k^T / tau, differentiable w.r.t. q and k (gcf_infonce_fwd / _bwd).
    The B x N logits exist only as tensor-core tiles in TMEM; the backward recomputes them tile by tile."""

    @staticmethod
    def forward(ctx, q, k, tau, cos, pos_idx, want_col):
        q, _ = _rows_view(q, "q")
        k, _ = _rows_view(k, "k")
        row, col, pos = infonce_stats_raw(q, k, tau, cos=cos, pos_idx=pos_idx, want_col=want_col)
        ctx.save_for_backward(q, k, row, col if want_col else None, pos_idx)
        ctx.tau, ctx.cos, ctx.want_col = float(tau), bool(cos), bool(want_col)
        if not want_col:
            col = row.new_zeros(0)
        ctx.set_materialize_grads(False)
        return row, col, pos

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_row, g_col, g_pos):
        lib = _lib.load()
        q, k, row, col, pos_idx = ctx.saved_tensors
        if not ctx.want_col:
            g_col = None
        if g_row is None and g_col is None and g_pos is None:
            return (None,) * 6
        m, d = q.shape
        n = k.shape[0]
        dev = q.device
        f = lambda t: None if t is None else t.contiguous().to(torch.float32)
        g_row, g_col, g_pos = f(g_row), f(g_col), f(g_pos)
        need_q, need_k = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        gq = torch.empty_like(q) if need_q else None
        gk = torch.empty_like(k) if need_k else None
        ws_bytes = lib.gcf_infonce_workspace_bytes(m, n, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gcf_infonce_bwd(_lib.ptr(q), q.stride(0), m, _lib.ptr(k), k.stride(0), n, d, 1 if ctx.cos else 0, ctx.tau,
                                       _lib.ptr(pos_idx), _lib.ptr(row if g_row is not None else None),
                                       _lib.ptr(col if g_col is not None else None), _lib.ptr(g_row), _lib.ptr(g_col),
                                       _lib.ptr(g_pos), _lib.ptr(gq), d, _lib.ptr(gk), d, _lib.ptr(ws), ws_bytes,
                                       _lib.current_stream()), "gcf_infonce_bwd")
        return gq, gk, None, None, None, None


def infonce_stats(q: torch.Tensor, k: torch.Tensor, tau: float, *, cos: bool = True, pos_idx=None, want_col: bool = False):
    """Differentiable (row_lse[M], col_lse[N] | None, pos[M]); every InfoNCE-family loss below is a few vector ops on these."""
    if pos_idx is not None:
        pos_idx = _idx(pos_idx, q.device, "pos_idx")
    row, col, pos = _InfoNCEStats.apply(q, k, float(tau), bool(cos), pos_idx, bool(want_col))
    return row, (col if want_col else None), pos


def info_nce(view1: torch.Tensor, view2: torch.Tensor, temperature: float, b_cos: bool = True) -> torch.Tensor:
    """InfoNCE(view1, view2, temperature, b_cos) of ncl.py:125-130 / ssl4rec.py:19-23:
    -mean(diag(log_softmax(v1^ v2^T / tau, dim=1)))."""
    row, _, pos = infonce_stats(view1, view2, temperature, cos=b_cos)
    return (row - pos).mean()


def ssl_layer_side(context_rows: torch.Tensor, all_rows: torch.Tensor, idx, tau: float) -> torch.Tensor:
    """One side of NCLModel.ssl_layer_loss (ncl.py:358-367): sum_b [ log sum_{all} exp(c^_b z^_./tau) - c^_b z^_{idx_b}/tau ]
    with the denominator over ALL rows of `all_rows` (B x U or B x I logits, never materialised)."""
    row, _, pos = infonce_stats(context_rows, all_rows, tau, cos=True, pos_idx=idx)
    return (row - pos).sum()


def batch_softmax(user_emb: torch.Tensor, item_emb: torch.Tensor, temperature: float) -> torch.Tensor:
    """batch_softmax_loss of ssl4rec.py:25-30: -mean log( exp(s_ii) / sum_j exp(s_ij) + 1e-6 )."""
    row, _, pos = infonce_stats(user_emb, item_emb, temperature, cos=True)
    return -torch.log(torch.exp(pos - row) + 1e-6).mean()


def info_nce_symmetric(z1: torch.Tensor, z2: torch.Tensor, temp: float = 0.2) -> torch.Tensor:
    """info_nce_loss of gcl.py:28-35: (CE(S, arange) + CE(S^T, arange)) / 2 over the full N x N logits."""
    row, col, pos = infonce_stats(z1, z2, temp, cos=True, want_col=True)
    return ((row - pos).mean() + (col - pos).mean()) / 2


class _DirectAU(torch.autograd.Function):
    """(align, unif(x), unif(y)) of directau.py:240-251 (gcf_directau_fwd / _bwd; Gram matrices on tensor cores)."""

    @staticmethod
    def forward(ctx, x, y, t):
        lib = _lib.load()
        x, ldx = _rows_view(x, "x")
        y, ldy = _rows_view(y, "y")
        b, d = x.shape
        if y.shape != x.shape:
            raise ValueError("x and y must have the same shape")
        out3 = torch.empty(3, dtype=torch.float32, device=x.device)
        ws_bytes = lib.gcf_directau_workspace_bytes(b, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.gcf_directau_fwd(_lib.ptr(x), ldx, _lib.ptr(y), ldy, b, d, float(t), _lib.ptr(out3), _lib.ptr(ws), ws_bytes,
                                        _lib.current_stream()), "gcf_directau_fwd")
        ctx.save_for_backward(x, y, out3)
        ctx.t = float(t)
        return out3

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g3):
        lib = _lib.load()
        x, y, out3 = ctx.saved_tensors
        b, d = x.shape
        g3 = g3.contiguous().to(torch.float32)
        gx, gy = torch.empty_like(x), torch.empty_like(y)
        ws_bytes = lib.gcf_directau_workspace_bytes(b, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        _lib.check(lib.gcf_directau_bwd(_lib.ptr(x), x.stride(0), _lib.ptr(y), y.stride(0), b, d, ctx.t, _lib.ptr(out3), _lib.ptr(g3),
                                        _lib.ptr(gx), d, _lib.ptr(gy), d, _lib.ptr(ws), ws_bytes, _lib.current_stream()),
                   "gcf_directau_bwd")
        return gx, gy, None


def directau_terms(x: torch.Tensor, y: torch.Tensor, t: float = 2.0) -> torch.Tensor:
    """[alignment(x, y), uniformity(x), uniformity(y)] (directau.py:245-251) in one fused evaluation.
    Fewer than two rows: uniformity is 0 by the reference's `pdist.numel() > 0` guard."""
    if x.shape[0] < 2:
        xn = torch.nn.functional.normalize(x, dim=-1); yn = torch.nn.functional.normalize(y, dim=-1)
        align = (xn - yn).pow(2).sum(1).mean()
        zero = align.new_zeros(())
        return torch.stack([align, zero, zero])
    return _DirectAU.apply(x, y, float(t))


def l2_reg_loss(reg: float, *args: torch.Tensor) -> torch.Tensor:
    """ncl.py:122-123 / directau.py:35 / ssl4rec.py:16: reg * sum_x |x|_F / rows(x)   (norm NOT squared)."""
    emb_loss = 0
    for emb in args:
        emb_loss = emb_loss + torch.norm(emb, p=2) / emb.shape[0]
    return emb_loss * reg


# =========================================================================================
# SpMM with the row-L2-normalise epilogue (sept.py:220-226, mhcn.py:440-457)
# =========================================================================================
class _SpMMNormalize(torch.autograd.Function):
    """(t, y) = (A @ x, F.normalize(A @ x, dim=1)): ONE launch writes both (fused epilogue); the backward folds the chain rule
    of the normalisation into the gradient of t and runs a single transposed SpMM."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, graph: CSRGraph):
        x = _f32c(x, "x")
        if x.shape[0] != graph.n_cols:
            raise ValueError(f"spmm_normalize: X has {x.shape[0]} rows, operator has {graph.n_cols} columns")
        t = torch.empty(graph.n_rows, x.shape[1], dtype=torch.float32, device=x.device)
        y = torch.empty_like(t)
        spmm_raw(graph, x, y=t, out=y, epilogue=_lib.EPILOGUE_L2NORM)
        ctx.graph = graph
        ctx.save_for_backward(t, y)
        ctx.set_materialize_grads(False)
        return t, y

    @staticmethod
    def backward(ctx, g_t, g_y):
        t, y = ctx.saved_tensors
        if g_t is None and g_y is None:
            return None, None
        total = None if g_t is None else _f32c(g_t, "grad")
        if g_y is not None:
            g_y = _f32c(g_y, "grad")
            norm = t.norm(dim=1, keepdim=True)
            through = (g_y - y * (y * g_y).sum(dim=1, keepdim=True)) / norm.clamp_min(1e-12)   # F.normalize eps
            through = torch.where(norm < 1e-12, g_y / 1e-12, through)                          # |t| below eps: y = t / eps
            total = through if total is None else total + through
        graph_t = ctx.graph.transpose()
        gx = torch.empty(graph_t.n_rows, total.shape[1], dtype=torch.float32, device=total.device)
        spmm_raw(graph_t, total.contiguous(), y=gx)
        return gx, None


def spmm_and_normalize(graph: CSRGraph, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(A @ x, F.normalize(A @ x, dim=1)) -- mhcn.py:440-457 keeps the raw product for the next layer and the normalised one
    for the layer sum."""
    return _SpMMNormalize.apply(x, graph)


def spmm_normalize(graph: CSRGraph, x: torch.Tensor) -> torch.Tensor:
    """F.normalize(A @ x, dim=1) (sept.py:222-224)."""
    return _SpMMNormalize.apply(x, graph)[1]
