"""LightGCN training over G GPUs of one NVSwitch box (north star "2/4/8-GPU runs"): the row-sharded
layout the north star describes (ShardedLightGCNTrainer, below) and a feature-sharded layout that needs no collective
in the propagation at all (FeatureShardedLightGCNTrainer, end of file; measured comparison in DESIGN.md section 6).

One process per GPU (torch.distributed, NCCL).  The reference has no distributed code at all; this is the
scaling path of the same step as lightgcn.FusedLightGCNTrainer:

  * node v (user u -> v = u, item i -> v = U + i) is owned by rank v % G at local row v // G (cyclic: the
    power-law hubs, which have the lowest ids in the synthetic graphs and arbitrary ids in real data, spread
    evenly).  Rank r keeps rows {v : v % G == r} of the embedding table, of the Adam state and of the
    normalised adjacency; column ids of its CSR block are "gathered positions" pos(v) = (v % G) * n_loc + v // G,
    i.e. the row of v inside an all-gathered [G * n_loc, d] buffer.
  * forward : per layer ONE all-gather of the [n_loc, d] fp32 shards (in place, the local SpMM writes straight
    into the rank's slot of the next gather buffer), then the local row-block SpMM; the layer sum is folded into
    the last SpMM's epilogue; one more all-gather publishes the final embeddings for the loss.
  * loss    : training triples are sharded by edge (E/G per rank); fused gather+BPR on the gathered final
    embeddings; the [G * n_loc, d] gradient partials are combined with ONE reduce-scatter.
  * backward: A is symmetric, so rank r's row block applied to the all-gathered G(k+1) yields its rows of
    A^T G(k+1): K more all-gathers.  Adam runs on the local shard.

Communication volume per step: (2K+1) all-gathers + 1 reduce-scatter of N*d*4 bytes each (3.84 GB at cfg 5);
the data path has no other collective.  ShardPlan is pure index arithmetic (CPU-testable with gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib, peer
from .graph import CSRGraph
from .tables import xavier_uniform_table


@dataclass(frozen=True)
class ShardPlan:
    """Cyclic node partition over `world` ranks."""

    n_users: int
    n_items: int
    world: int

    @property
    def n_nodes(self) -> int:
        return self.n_users + self.n_items

    @property
    def n_loc(self) -> int:
        """Rows per rank (the last ranks may carry one padding row)."""
        return -(-self.n_nodes // self.world)

    @property
    def n_padded(self) -> int:
        return self.n_loc * self.world

    def owner(self, v):
        return v % self.world

    def local_row(self, v):
        return v // self.world

    def gathered_pos(self, v):
        """Row of node v inside an all-gathered [world * n_loc, d] buffer."""
        return (v % self.world) * self.n_loc + v // self.world

    def node_of_pos(self, pos):
        """Inverse of gathered_pos (padding rows map to ids >= n_nodes)."""
        return (pos % self.n_loc) * self.world + pos // self.n_loc

    def local_nodes(self, rank: int, device=None) -> torch.Tensor:
        return torch.arange(rank, self.n_nodes, self.world, dtype=torch.int64, device=device)

    def local_block_coo(self, users: torch.Tensor, items: torch.Tensor, rank: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Rows of the bidirectional adjacency [[0,R],[R^T,0]] owned by `rank`:
        (local row ids, gathered-position column ids), duplicates kept."""
        u = users.to(torch.int64)
        i = items.to(torch.int64) + self.n_users
        mu = (u % self.world) == rank          # user rows owned here: entries (u, i)
        mi = (i % self.world) == rank          # item rows owned here: entries (i, u)
        rows = torch.cat([u[mu] // self.world, i[mi] // self.world])
        cols = torch.cat([self.gathered_pos(i[mu]), self.gathered_pos(u[mi])])
        return rows, cols

    def triple_range(self, n_triples: int, rank: int) -> Tuple[int, int]:
        """Contiguous slice of the training triples handled by `rank` (balanced to within one)."""
        base, rem = divmod(n_triples, self.world)
        lo = rank * base + min(rank, rem)
        return lo, lo + base + (1 if rank < rem else 0)


def shard_table(full: torch.Tensor, plan: ShardPlan, rank: int) -> torch.Tensor:
    """[N, d] -> this rank's [n_loc, d] rows (zero padded)."""
    out = torch.zeros(plan.n_loc, full.shape[1], dtype=full.dtype, device=full.device)
    rows = full[rank::plan.world]
    out[: rows.shape[0]] = rows
    return out


def unshard_table(gathered: torch.Tensor, plan: ShardPlan) -> torch.Tensor:
    """All-gathered [world * n_loc, d] buffer -> [N, d] in node order."""
    pos = plan.gathered_pos(torch.arange(plan.n_nodes, dtype=torch.int64, device=gathered.device))
    return gathered[pos]


class ShardedLightGCNTrainer:
    """Full-batch LightGCN step (lightgcn.py:83-120), row-sharded over the default process group."""

    def __init__(self, users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int, *, d: int = 64,
                 n_layers: int = 3, lr: float = 0.01, reg_weight: float = 1e-4, seed: int = 0,
                 init_table: Optional[torch.Tensor] = None, feature_shards: int = 1):
        """feature_shards = F > 1 gives the 2-D layout: the G ranks form R = G / F row groups x F feature groups; a rank
        owns N/R rows x d/F columns.  All-gathers / reduce-scatters run inside a row group (ranks with the same feature
        slice) on d/F-wide rows -- F times less volume -- and the BPR scores are completed by one all-reduce of E/R floats
        inside the feature group (see FeatureShardedLightGCNTrainer for that half)."""
        if not dist.is_initialized():
            raise RuntimeError("ShardedLightGCNTrainer needs an initialised torch.distributed process group")
        self.lib = _lib.load()
        g_rank, g_world = dist.get_rank(), dist.get_world_size()
        if feature_shards < 1 or g_world % feature_shards != 0:
            raise ValueError(f"feature_shards={feature_shards} must divide the world size {g_world}")
        self.fs = F = feature_shards
        self.world = R = g_world // F          # size of a row group: the sharding factor of the node dimension
        self.rank = g_rank // F                # this rank's row shard
        self.feat_rank = g_rank % F
        self.row_group = self.feat_group = None
        if F > 1:  # every rank creates every group, in the same order
            for f in range(F):
                grp = dist.new_group([r * F + f for r in range(R)])
                if f == self.feat_rank:
                    self.row_group = grp
            for r in range(R):
                grp = dist.new_group([r * F + f for f in range(F)])
                if r == self.rank:
                    self.feat_group = grp
        d_full = d
        lo_c, hi_c = feature_slice(d_full, F, self.feat_rank)
        d = hi_c - lo_c                        # local width: everything below works on [*, d/F] matrices
        dev = users.device
        if not users.is_cuda:
            raise RuntimeError("ShardedLightGCNTrainer: tensors must be on the rank's CUDA device")
        self.dev = dev
        self.plan = plan = ShardPlan(n_users, n_items, self.world)
        self.n_users, self.n_items, self.d, self.k = n_users, n_items, d, n_layers
        self.lr, self.reg, self.seed = lr, reg_weight, seed
        n_loc, n_pad = plan.n_loc, plan.n_padded
        self.rows_per_rank = n_loc
        self.n_edges = int(users.numel())

        # ---- local row block of D^-1/2 A D^-1/2 (values from the GLOBAL degree vector) ----
        lib, st = self.lib, _lib.current_stream()
        ei_rows = torch.cat([users, items + n_users]).contiguous()
        deg = torch.empty(plan.n_nodes, dtype=torch.int32, device=dev)
        _lib.check(lib.gcf_degree_count(_lib.ptr(ei_rows), ei_rows.numel(), _lib.ptr(deg), plan.n_nodes, st), "gcf_degree_count")
        del ei_rows
        dinv = deg.to(torch.float32).sqrt_().reciprocal_()
        dinv[torch.isinf(dinv)] = 0.0                       # inf -> 0 (selfcf.py:246)
        dinv_pos = torch.zeros(n_pad, dtype=torch.float32, device=dev)   # indexed by gathered position
        dinv_pos[plan.gathered_pos(torch.arange(plan.n_nodes, device=dev))] = dinv
        rows, cols = plan.local_block_coo(users, items, self.rank)
        block = CSRGraph.from_coo(rows, cols, None, n_loc, n_pad, norm="none")
        del rows, cols
        row_scale = dinv_pos[self.rank * n_loc:(self.rank + 1) * n_loc].contiguous()
        _lib.check(lib.gcf_scale_csr_values(_lib.ptr(block.row_ptr), _lib.ptr(block.col_idx), _lib.ptr(block.vals), n_loc,
                                            _lib.ptr(row_scale), _lib.ptr(dinv_pos), _lib.ptr(block.vals), st),
                   "gcf_scale_csr_values")
        self.block = block
        self.local_nnz = block.nnz
        self.ws, self.ws_bytes = block.workspace(d)

        # ---- parameters / optimiser state: local rows only ----
        if init_table is not None:
            self.table = shard_table(init_table.to(dev)[:, lo_c:hi_c].contiguous(), plan, self.rank).contiguous()
        else:
            # a pure function of (seed, node, column): independent of the layout and of the number of ranks
            full = xavier_uniform_table(n_users, n_items, d_full, seed=seed, device=dev, cols=(lo_c, hi_c))
            self.table = shard_table(full, plan, self.rank).contiguous()
            del full
        self.exp_avg = torch.zeros_like(self.table)
        self.exp_avg_sq = torch.zeros_like(self.table)

        # ---- gather buffers: E0..E(K-1) gathered, final gathered, gradient gathered ----
        self.full = [torch.zeros(n_pad, d, device=dev) for _ in range(n_layers)]
        self.final_full = torch.zeros(n_pad, d, device=dev)
        self.g_full = torch.zeros(n_pad, d, device=dev)
        self.g_loc = torch.zeros(n_loc, d, device=dev)
        self.gk_full = [torch.zeros(n_pad, d, device=dev) for _ in range(2)]
        self.g_x0 = torch.zeros(n_loc, d, device=dev)

        # ---- this rank's triples, in gathered-position space ----
        # The canonical numbering of the triples is the user-major order of the whole list (the one FusedLightGCNTrainer and
        # the feature-sharded trainer use): rank r takes positions [lo, hi) of it, so the Philox negatives -- keyed by that
        # position -- and with them the whole optimisation trajectory do not depend on the layout or the number of ranks.
        u64, i64 = users.to(torch.int64), items.to(torch.int64)
        order_global = torch.argsort(u64 * n_items + i64) if self.n_edges > 1 else torch.arange(self.n_edges, device=dev)
        lo, hi = plan.triple_range(self.n_edges, self.rank)
        self.n_triples = hi - lo
        sel = order_global[lo:hi]
        del order_global
        pos_u = plan.gathered_pos(u64[sel])
        pos_i = plan.gathered_pos(i64[sel] + n_users)
        del u64, i64
        # user-major order in gathered-position space inside the slice: the fused BPR kernel keeps the user row / gradient in
        # registers over a run (the order of the triples is free: the loss is a mean over all of them)
        self.order = torch.argsort(pos_u * plan.n_padded + pos_i) if self.n_triples > 1 else None
        if self.order is not None:
            pos_u, pos_i, sel = pos_u[self.order], pos_i[self.order], sel[self.order]
        self.sel = sel.contiguous()            # position in the caller's list of each local triple (externally drawn negatives)
        self.pos_u, self.pos_i = pos_u.contiguous(), pos_i.contiguous()
        self.neg_raw = torch.empty(max(self.n_triples, 1), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        self.bpr_ws_bytes = lib.gcf_bpr_workspace_bytes(self.n_triples)
        self.bpr_ws = torch.empty(self.bpr_ws_bytes, dtype=torch.uint8, device=dev)
        self.triple_offset = lo
        self.step_count = 0
        self.scores = torch.empty(max(self.n_triples, 1), dtype=torch.float32, device=dev) if F > 1 else None
        self.coef = torch.empty(max(self.n_triples, 1), dtype=torch.float32, device=dev) if F > 1 else None
        self.loss_reg = torch.zeros((), dtype=torch.float32, device=dev)
        self.launches_per_step = 2 * n_layers + (4 if F == 1 else 7)
        self.collectives_per_step = 2 * n_layers + 2 + (1 if F > 1 else 0)

    # -------------------------------------------------------------------------------------------------
    def _slot(self, buf: torch.Tensor) -> torch.Tensor:
        return buf[self.rank * self.plan.n_loc:(self.rank + 1) * self.plan.n_loc]

    def _spmm(self, x_full, y, out, alpha, post, addends, betas):
        lib = self.lib
        _lib.check(lib.gcf_spmm_csr_f32(self.block.struct_ref(), self.d, _lib.ptr(x_full), self.d,
                                        _lib.ptr(y), self.d if y is not None else 0,
                                        _lib.ptr(out), self.d if out is not None else 0,
                                        _lib.EPILOGUE_NONE, alpha, post, len(addends), _lib.ptr_array(list(addends)),
                                        _lib.float_array(list(betas)), _lib.ptr(self.ws), self.ws_bytes, 0,
                                        _lib.current_stream()), "gcf_spmm_csr_f32")

    def index_buffers(self) -> List[torch.Tensor]:
        """The device tensors that hold this rank's training triples -- the per-step INPUT of the step (bench.py's e2e arm
        refills them from pinned host memory every step)."""
        return [self.pos_u, self.pos_i]

    def step(self, neg_items: Optional[torch.Tensor] = None, marks: Optional[list] = None, wait_before_loss=None) -> torch.Tensor:
        """One optimisation step; returns the (global) loss.  wait_before_loss: optional CUDA event the loss phase waits for
        (an asynchronous refill of index_buffers()).  neg_items: optional pre-drawn raw item ids for ALL triples in
        the caller's order, identical on every rank (parity tests); otherwise Philox negatives keyed by the GLOBAL triple index (gcf_sample_negatives_at),
        so the draw is independent of the number of ranks."""
        lib, st, plan, d, K = self.lib, _lib.current_stream(), self.plan, self.d, self.k
        self.step_count += 1
        ev = (lambda: torch.cuda.Event(enable_timing=True)) if marks is not None else None
        spmm_marks: List = []

        def timed_spmm(*a):
            if ev is None:
                return self._spmm(*a)
            e0, e1 = ev(), ev()
            e0.record(); self._spmm(*a); e1.record()
            spmm_marks.append((e0, e1, 1))

        # ---- forward: E0 gathered, then K SpMM with K-1 more gathers ----
        rg = self.row_group
        self._slot(self.full[0]).copy_(self.table)
        dist.all_gather_into_tensor(self.full[0], self._slot(self.full[0]), group=rg)
        final_loc = self._slot(self.final_full)
        for k in range(1, K + 1):
            if k < K:
                y = self._slot(self.full[k])
                timed_spmm(self.full[k - 1], y, None, 1.0, 1.0, [], [])
                dist.all_gather_into_tensor(self.full[k], y, group=rg)
            else:  # last layer: final = sum_k E(k) (lightgcn.py:26), E(K) itself is not stored
                adds = [self._slot(self.full[q]) for q in range(K)]
                timed_spmm(self.full[K - 1], None, final_loc, 1.0, 1.0, adds, [1.0] * K)
        dist.all_gather_into_tensor(self.final_full, final_loc, group=rg)

        # ---- loss on this rank's triples ----
        if wait_before_loss is not None:
            torch.cuda.current_stream().wait_event(wait_before_loss)
        if neg_items is None:
            # Philox slot = position of the triple in the caller's (global) list: rank r draws the window [lo, hi) of the
            # stream a single-GPU run over the same list draws, then applies its local user-major permutation
            _lib.check(lib.gcf_sample_negatives_at(self.seed, self.step_count, self.triple_offset, None, self.n_triples, 1,
                                                   self.n_items, None, None, 1, _lib.ptr(self.neg_raw), st),
                       "gcf_sample_negatives_at")
            if self.order is not None and self.n_triples > 0:
                self.neg_raw[: self.n_triples] = self.neg_raw[: self.n_triples][self.order]
            neg_raw = self.neg_raw[: self.n_triples]
        else:
            neg_raw = neg_items.to(torch.int64).reshape(-1)[self.sel]
        neg = plan.gathered_pos(neg_raw + self.n_users).contiguous()
        w = 1.0 / self.n_edges
        # per-rank partial of the global mean: reduction = sum, loss and gradients scaled by 1/E afterwards / inside
        self.g_full.zero_()
        tables = (_lib.ptr(self.final_full), d, _lib.ptr(self.final_full), d, d, _lib.ptr(self.pos_u), _lib.ptr(self.pos_i),
                  _lib.ptr(neg), self.n_triples, 1)
        if self.fs == 1:
            # forward and backward in one pass over the triples (rows gathered once)
            _lib.check(lib.gcf_bpr_fwd_bwd(*tables, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_SUM, self.reg / w, self.reg / w, 0.0, w,
                                           _lib.ptr(self.loss), None, _lib.ptr(self.g_full), d, _lib.ptr(self.g_full), d,
                                           _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st), "gcf_bpr_fwd_bwd")
            loss_local = self.loss * w
        else:
            # partial scores on the local columns -> all-reduce inside the feature group -> loss -> local gradients
            _lib.check(lib.gcf_bpr_fwd(*tables, _lib.BPR_RAW_SCORE, 0.0, _lib.REDUCE_SUM, self.reg / w, self.reg / w, 0.0,
                                       _lib.ptr(self.loss_reg), _lib.ptr(self.scores), _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st),
                       "gcf_bpr_fwd")
            dist.all_reduce(self.scores, op=dist.ReduceOp.SUM, group=self.feat_group)
            _lib.check(lib.gcf_bpr_coef_from_scores(_lib.ptr(self.scores), self.n_triples, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_SUM,
                                                    _lib.ptr(self.loss), _lib.ptr(self.coef), _lib.ptr(self.bpr_ws),
                                                    self.bpr_ws_bytes, st), "gcf_bpr_coef_from_scores")
            gscale = torch.full((), w, dtype=torch.float32, device=self.dev)
            _lib.check(lib.gcf_bpr_bwd(*tables, _lib.ptr(self.coef), _lib.ptr(gscale), self.reg / w, self.reg / w, 0.0,
                                       _lib.ptr(self.g_full), d, _lib.ptr(self.g_full), d, st), "gcf_bpr_bwd")
            # the pointwise part is replicated inside the feature group: count it once
            loss_local = (self.loss / self.fs + self.loss_reg) * w
        dist.reduce_scatter_tensor(self.g_loc, self.g_full, op=dist.ReduceOp.SUM, group=rg)

        # ---- backward propagation: G(k) = A G(k+1) + g,  G(K) = g  (scale 1: 'sum' combination) ----
        cur_full = self.gk_full[0]
        self._slot(cur_full).copy_(self.g_loc)
        dist.all_gather_into_tensor(cur_full, self._slot(cur_full), group=rg)
        which = 1
        for k in range(K - 1, -1, -1):
            if k > 0:
                nxt = self.gk_full[which]
                out = self._slot(nxt)
                timed_spmm(cur_full, None, out, 1.0, 1.0, [self.g_loc], [1.0])
                dist.all_gather_into_tensor(nxt, out, group=rg)
                cur_full = nxt
                which ^= 1
            else:
                timed_spmm(cur_full, None, self.g_x0, 1.0, 1.0, [self.g_loc], [1.0])

        _lib.check(lib.gcf_adam_step(_lib.ptr(self.table), _lib.ptr(self.g_x0), _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq),
                                     self.table.numel(), self.lr, 0.9, 0.999, 1e-8, 0.0, 0, self.step_count, st), "gcf_adam_step")
        loss = loss_local.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        if marks is not None:
            marks.extend(spmm_marks)
        return loss

    def gathered_table(self) -> torch.Tensor:
        """[N, d] table in node order on every rank (for checks / evaluation)."""
        buf = torch.zeros(self.plan.n_padded, self.d, device=self.dev)
        self._slot(buf).copy_(self.table)
        dist.all_gather_into_tensor(buf, self._slot(buf), group=self.row_group)
        rows = unshard_table(buf, self.plan).contiguous()
        if self.fs == 1:
            return rows
        parts = [torch.empty_like(rows) for _ in range(self.fs)]
        dist.all_gather(parts, rows, group=self.feat_group)
        return torch.cat(parts, dim=1)


# =====================================================================================================
# Feature-sharded layout: every rank holds ALL N rows but only d/G columns of every [N, d] matrix.
# =====================================================================================================
def feature_slice(d: int, world: int, rank: int) -> Tuple[int, int]:
    """Columns [lo, hi) of the embedding dimension owned by `rank` (d must split into multiples of 4 floats)."""
    if d % world != 0 or (d // world) % 4 != 0:
        raise ValueError(f"feature sharding needs d/G to be a multiple of 4 (d={d}, G={world})")
    dg = d // world
    return rank * dg, (rank + 1) * dg


@dataclass(frozen=True)
class UserBlockPlan:
    """Contiguous user blocks with balanced numbers of training triples (pure index arithmetic, CPU-testable).

    The triples are kept user-major; rank r evaluates the loss on the triples of users [cuts[r], cuts[r+1]), i.e. on
    positions [triple_cuts[r], triple_cuts[r+1]) of the sorted list.  Cuts fall on user boundaries, so a user's row and
    its gradient live on exactly one rank."""

    cuts: Tuple[int, ...]
    triple_cuts: Tuple[int, ...]

    @property
    def world(self) -> int:
        return len(self.cuts) - 1

    def block(self, rank: int) -> Tuple[int, int]:
        return self.cuts[rank], self.cuts[rank + 1]

    def block_sizes(self) -> List[int]:
        return [self.cuts[r + 1] - self.cuts[r] for r in range(self.world)]

    def triple_range(self, rank: int) -> Tuple[int, int]:
        return self.triple_cuts[rank], self.triple_cuts[rank + 1]

    @staticmethod
    def build(sorted_users: torch.Tensor, n_users: int, world: int) -> "UserBlockPlan":
        """sorted_users: the user id of every triple, ascending (int64, any device)."""
        e = int(sorted_users.numel())
        cuts, tcuts = [0], [0]
        for r in range(1, world):
            if e == 0:
                cuts.append(cuts[-1]); tcuts.append(0)
                continue
            u = int(sorted_users[min(r * e // world, e - 1)].item())   # the user that owns the ideal cut position ...
            u = max(u, cuts[-1])
            t = int(torch.searchsorted(sorted_users, torch.tensor([u], dtype=sorted_users.dtype, device=sorted_users.device))[0].item())
            cuts.append(u); tcuts.append(t)                            # ... starts the next block with ALL its triples
        cuts.append(n_users); tcuts.append(e)
        return UserBlockPlan(tuple(cuts), tuple(tcuts))


def owner_major_row(u: torch.Tensor, world: int, rows_per_owner: int) -> torch.Tensor:
    """Row of user u in the owner-major numbering of the feature-sharded tables: the users of rank g = u % world occupy the
    contiguous block [g * rows_per_owner, (g + 1) * rows_per_owner), in ascending user order (rows_per_owner = ceil(U / world);
    blocks of ranks with one user fewer end in a padding row)."""
    return (u % world) * rows_per_owner + torch.div(u, world, rounding_mode="floor")


def cyclic_user_shard(sorted_users: torch.Tensor, sorted_items: torch.Tensor, n_users: int, world: int, rank: int):
    """Users dealt out cyclically over the ranks (user u belongs to rank u % world, where it is row u // world of the rank's
    user set): pure index arithmetic, any device (CPU-testable).

    sorted_users / sorted_items: the training triples in the global user-major order.  Returns (positions, loc_u, loc_i,
    rows_per_rank): the positions of this rank's triples in the global list (ascending, so the rank's list is user-major too and
    its Philox negatives can be drawn by position), their user rows inside the rank's user set, their items, and the number of
    users of every rank."""
    mine = (sorted_users % world) == rank
    positions = torch.nonzero(mine).reshape(-1)
    loc_u = torch.div(sorted_users[mine], world, rounding_mode="floor").contiguous()
    loc_i = sorted_items[mine].contiguous()
    rows_per_rank = [max(0, (n_users - g + world - 1) // world) for g in range(world)]
    return positions, loc_u, loc_i, rows_per_rank


class FeatureShardedLightGCNTrainer:
    """The same full-batch step, parallelised over the embedding dimension.

    The propagation is linear and acts on every feature column independently: E(k+1)[:, c] = A E(k)[:, c].  A rank
    that owns columns [lo, hi) of the tables therefore runs all K layers forward and backward, and Adam, on its
    [N, d/G] slice WITHOUT any exchange -- against (2K+1) all-gathers + 1 reduce-scatter of N*d*4 bytes in the
    row-sharded layout.  The only coupling is the BPR score <u, p - n>, a sum over all d columns:

        partial scores on the local columns (gcf_bpr_fwd, GCF_BPR_RAW_SCORE)      [E floats]
        -> ONE all-reduce of E floats per step (0.4 GB at cfg 5, vs 30.7 GB above)
        -> loss / dl/dx on the complete scores (gcf_bpr_coef_from_scores), identical on every rank
        -> gradients w.r.t. the local columns (gcf_bpr_bwd), transpose propagation, Adam: local.

    Costs: the CSR (int32 structure + fp32 values, 1.6 GB at cfg 5) and the triple index arrays are replicated, every
    rank walks all nnz / all triples (with d/G-wide rows: 32 B at d = 64, G = 8), and the three row gathers of the loss
    are done twice (score pass, gradient pass).  Negatives are drawn with the same Philox (seed, step) on every rank.
    That is loss_layout="scores".

    loss_layout="rows" (default) keeps the propagation as above but evaluates the loss on FULL-width rows, each rank on
    the triples of one contiguous user block (UserBlockPlan), with the single-GPU fused forward+backward BPR kernel:

        item slices  [I, d/G]  --all-gather-->        [G, I, d/G]  --gcf_slices_to_rows--> item table [I, d]
        user slices  [U, d/G]  --all-to-all(blocks)-> [G, Ub, d/G] --gcf_slices_to_rows--> my users   [Ub, d]
        gcf_bpr_fwd_bwd on E/G triples (rows gathered once, gradients produced in the same pass)
        item grads [I, d]  --gcf_rows_to_slices--> [G, I, d/G]  --reduce-scatter--> my columns of dL/d(final items)
        user grads [Ub, d] --gcf_rows_to_slices--> [G, Ub, d/G] --all-to-all-->     my columns of dL/d(final users)

    2 x (I + U/G) x d x 4 x (G-1)/G bytes cross the links per rank and step (2.8 GB at cfg 5, G = 8 -- against 7 x 3.4 GB
    in the row-sharded layout) and every rank gathers 3 x E/G full rows once instead of 3 x E narrow rows twice: measured
    one-rank cost of the loss part at d/G = 8 in the "scores" layout 17.9 ms of a 44.5 ms step (profiles/r01_exp_l2_window.md).
    """

    def __init__(self, users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int, *, d: int = 64,
                 n_layers: int = 3, lr: float = 0.01, reg_weight: float = 1e-4, seed: int = 0,
                 init_table: Optional[torch.Tensor] = None, loss_layout: str = "rows", overlap: bool = False,
                 exchange: str = "peer", user_rows: str = "owner"):
        """exchange (loss_layout="rows"): "peer" (default) -- the column slices are read straight out of the peers' memory
        over NVLink by one kernel per direction that also converts the layout (csrc/peer.cu; two device-side barriers over peer
        flag words per step; the only NCCL call left is the all-reduce of the scalar loss); "nccl" -- the r01 path: all-gather / all-to-all / reduce-scatter with a
        layout pass on either side.
        user_rows (exchange="peer"): "owner" (default) -- inside the trainer the user rows of every [N, d/G] table are stored
        owner-major (user u, dealt to rank u % G, sits at row (u % G) * ceil(U / G) + u // G; the graph is built on that
        numbering), so a rank's users are ONE contiguous block of every peer's slice and both user transfers are contiguous on
        the remote side (the links move 32-byte pieces of wider remote rows at half the rate); "natural" keeps node order.
        overlap (loss_layout="rows", n_layers >= 2): the last forward layer and the first backward layer are launched as
        an item-row block and a user-row block of the (bipartite) operator, so that the item all-gather runs while the user
        rows are still being computed and the item reduce-scatter while the item rows of the first backward product are.
        Measured neutral on cfg5 (2 GPUs 61.7 vs 61.4 ms, 4 GPUs 40.1 vs 40.2 ms: the collectives compete with the SpMM for
        HBM / L2 bandwidth), hence off by default."""
        if not dist.is_initialized():
            raise RuntimeError("FeatureShardedLightGCNTrainer needs an initialised torch.distributed process group")
        if loss_layout not in ("rows", "scores"):
            raise ValueError("loss_layout must be 'rows' or 'scores'")
        if exchange not in ("peer", "nccl"):
            raise ValueError("exchange must be 'peer' or 'nccl'")
        self.loss_layout = loss_layout
        if not users.is_cuda:
            raise RuntimeError("FeatureShardedLightGCNTrainer: tensors must be on the rank's CUDA device")
        self.exchange = exchange if loss_layout == "rows" else "nccl"
        if self.exchange == "peer" and not peer.available(users.device):
            self.exchange = "nccl"      # agreed on by all ranks (collective probe); reported by bench.py's `parallelism` string
        self.overlap = bool(overlap) and loss_layout == "rows" and n_layers >= 2
        self.lib = _lib.load()
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.lo, self.hi = feature_slice(d, self.world, self.rank)
        self.dg = self.hi - self.lo
        dev = users.device
        self.dev, self.d_full = dev, d
        if user_rows not in ("owner", "natural"):
            raise ValueError("user_rows must be 'owner' or 'natural'")
        self.true_users = n_users
        self.owner_major = user_rows == "owner" and self.exchange == "peer" and self.world > 1
        users_g, u_rows = users, n_users
        if self.owner_major:
            G = self.world
            self.ubm = -(-n_users // G)                       # rows reserved per owner (the last owners may have one fewer user)
            u_rows = self.ubm * G                             # user rows of the tables, padding rows included (isolated, all-zero)
            users_g = owner_major_row(users.to(torch.int64), G, self.ubm)
        # n_users = number of user ROWS in the [N, d/G] tables (what every row offset below needs); true_users = U
        self.n_users, self.n_items, self.n = u_rows, n_items, u_rows + n_items
        self.k, self.lr, self.reg, self.seed = n_layers, lr, reg_weight, seed
        self.graph = CSRGraph.from_pairs(users_g, items, u_rows, n_items, norm="sym")  # replicated operator
        del users_g
        self.local_nnz = self.graph.nnz
        self.rows_per_rank = self.n
        n, dg = self.n, self.dg
        if init_table is not None:
            self.table = init_table.to(dev)[:, self.lo:self.hi].contiguous()
        else:
            self.table = xavier_uniform_table(n_users, n_items, d, seed=seed, device=dev, cols=(self.lo, self.hi))
        if self.owner_major:
            nat = self.table
            self.table = torch.zeros(n, dg, device=dev)
            self.table[self._user_row(torch.arange(n_users, device=dev))] = nat[:n_users]
            self.table[u_rows:] = nat[n_users:]
            del nat
        pos_u, pos_i = users.to(torch.int64), items.to(torch.int64)
        self.n_edges = self.n_triples = int(pos_u.numel())
        self.order = torch.argsort(pos_u * n_items + pos_i) if self.n_triples > 1 else None  # user-major (see lightgcn.py)
        if self.order is not None:
            pos_u, pos_i = pos_u[self.order], pos_i[self.order]
        new = lambda: torch.empty(n, dg, device=dev)
        self.layers = [new() for _ in range(n_layers - 1)] + [None]
        self.final, self.g_final = new(), new()
        self.fused_adam = True   # bench.py: the backward timing pair includes the optimiser epilogue
        self.ping = new() if n_layers > 1 else None
        self.pong = new() if n_layers > 1 else None
        self.exp_avg, self.exp_avg_sq = torch.zeros(n, dg, device=dev), torch.zeros(n, dg, device=dev)
        self.loss_reg = torch.zeros((), dtype=torch.float32, device=dev)
        self.loss_pt = torch.zeros((), dtype=torch.float32, device=dev)
        self.ws, self.ws_bytes = self.graph.workspace(dg)
        self.step_count = 0
        if loss_layout == "scores":
            self.pos_u, self.pos_i = pos_u.contiguous(), pos_i.contiguous()
            self.n_local = self.n_triples
            self.neg = torch.empty(max(self.n_triples, 1), dtype=torch.int64, device=dev)
            self.scores = torch.empty(max(self.n_triples, 1), dtype=torch.float32, device=dev)
            self.coef = torch.empty(max(self.n_triples, 1), dtype=torch.float32, device=dev)
            self.bpr_ws_bytes = self.lib.gcf_bpr_workspace_bytes(self.n_triples)
            self.bpr_ws = torch.empty(self.bpr_ws_bytes, dtype=torch.uint8, device=dev)
            # K fwd + K bwd SpMM (Adam fused into the last), sampler, score pass + reduce, coef + reduce, gradient pass
            self.launches_per_step = 2 * n_layers + 6
            self.collectives_per_step = 2  # E-float score all-reduce + scalar loss all-reduce
        else:
            G = self.world
            self.item_full = torch.empty(n_items, d, device=dev)
            if self.exchange == "peer":
                # users are dealt out cyclically (u % G): triples AND user rows are balanced over the ranks whatever the degree
                # law (contiguous triple-balanced blocks give the hub block few users and the tail block many, so the row
                # exchange of the two directions is lopsided and every rank waits for the slowest one twice).  A peer pull
                # can read any stride; only NCCL's all-to-all needs contiguous blocks.
                r = self.rank
                self.positions, self.loc_u, self.loc_i, self.block_rows = cyclic_user_shard(pos_u, pos_i, n_users, G, r)
                del pos_u, pos_i
                self.n_local = int(self.loc_u.numel())
                self.ub = ub = self.block_rows[r]
                self.user_full = torch.empty(ub, d, device=dev)
                # what the peers read: my column slice of the final embeddings, my full-width gradient partials
                self._barrier = peer.PeerBarrier(dev)
                self._pb_final = peer.PeerBuffer(n, dg, dev)
                self._pb_gu = peer.PeerBuffer(max(ub, 1), d, dev)
                # item gradients travel the other way: every rank PUSHES the column pieces of its [I, d] partial into the
                # owners' staging slices [G][I][dg] (contiguous on the remote side: 32-byte pieces of a wider remote row
                # move at half the rate, tools/peer_bw.py), which the owner then sums locally in rank order
                self._pb_stage = peer.PeerBuffer(G * n_items, dg, dev)
                self.final = self._pb_final.tensor
                self.g_item_full = torch.empty(n_items, d, device=dev)
                self._rows_items64 = peer.int64_array([n_items] * G)
                self.g_user_full = self._pb_gu.tensor[:ub]
                self._block_rows64 = peer.int64_array(self.block_rows)
                self._user_dst_off64 = peer.int64_array([g * dg for g in range(G)])   # user u = g + j G -> row u of [U, dg]
                if self.owner_major:
                    # user gradients are PUSHED by their owner into the contiguous block of my slice that holds its users; the
                    # block starts zeroed and nobody ever writes its padding rows, so they stay zero (and so do their table rows)
                    self._pb_gfinal = peer.PeerBuffer(n, dg, dev)
                    self.g_final = self._pb_gfinal.tensor
                    self._rows_mine64 = peer.int64_array([ub] * G)
            else:
                self.blocks = UserBlockPlan.build(pos_u, n_users, G)
                self.block_rows = self.blocks.block_sizes()
                u_lo, u_hi = self.blocks.block(self.rank)
                t_lo, t_hi = self.blocks.triple_range(self.rank)
                self.t_lo, self.t_hi, self.ub = t_lo, t_hi, u_hi - u_lo
                self.n_local = t_hi - t_lo
                self.loc_u = (pos_u[t_lo:t_hi] - u_lo).contiguous()       # row inside my user block
                self.loc_i = pos_i[t_lo:t_hi].contiguous()
                del pos_u, pos_i
                ub = self.ub
                self.user_full = torch.empty(ub, d, device=dev)
                self.item_blk = torch.empty(G, n_items, dg, device=dev)    # all-gather target / reduce-scatter source
                self.g_item_full = torch.empty(n_items, d, device=dev)
                self.user_blk = torch.empty(G * ub, dg, device=dev)        # [G, Ub, d/G]: all-to-all target / source
                self.g_user_full = torch.empty(ub, d, device=dev)
            self.neg = torch.empty(max(self.n_local, 1), dtype=torch.int64, device=dev)
            self.bpr_ws_bytes = self.lib.gcf_bpr_workspace_bytes(self.n_local)
            self.bpr_ws = torch.empty(self.bpr_ws_bytes, dtype=torch.uint8, device=dev)
            # libgcf launches per step.  nccl: K fwd + K bwd SpMM (Adam fused into the last), sampler, 2 + 2 layout conversions,
            # fused BPR + its reduction.  peer: K + K SpMM, sampler, fused BPR + its reduction, 2 barriers, item gather, user
            # gather, item push, item sum, user push (or pull)
            self.launches_per_step = 2 * n_layers + (7 if self.exchange == "nccl" else 10)
            # nccl: item all-gather, user all-to-all, item reduce-scatter, user all-to-all, loss all-reduce; peer: two
            # stream-ordered barriers + the loss all-reduce (2 gathers + 1 sum + 1 copy replace the 4 layout passes)
            self.collectives_per_step = 5 if self.exchange == "nccl" else 3
            if self.overlap:
                # row blocks of the replicated operator as views of its arrays: user rows gather item columns only and
                # vice versa (bipartite), so each block depends on ONE side of the exchanged gradient / produces one side
                gr = self.graph
                cut = int(gr.row_ptr[self.n_users].item())

                def block(r0, r1, e0, e1):
                    rp = (gr.row_ptr[r0:r1 + 1] - e0).contiguous()
                    return CSRGraph(rp, gr.col_idx[e0:e1], gr.vals[e0:e1], r1 - r0, n, chunk=gr.chunk)
                self.g_users, self.g_items = block(0, self.n_users, 0, cut), block(self.n_users, n, cut, gr.nnz)
                self.ws_u, self.ws_u_bytes = self.g_users.workspace(dg)
                self.ws_i, self.ws_i_bytes = self.g_items.workspace(dg)
                self.p_buf = new()                      # A g_final, handed to the rest of the backward chain as extra[K-1]
                # high priority: the movers' few CTAs must get SM slots while a propagation launch has thousands queued
                self._side = torch.cuda.Stream(device=dev, priority=-1) if self.exchange == "peer" else None
                # two block launches instead of one (forward, backward) + the G(K-1) axpby; peer: two more barriers
                self.launches_per_step += 3 if self.exchange == "nccl" else 5

    def _user_row(self, u: torch.Tensor) -> torch.Tensor:
        """Table row of user u (owner-major numbering: owner block u % G, row u // G inside it)."""
        return owner_major_row(u, self.world, self.ubm) if self.owner_major else u

    def _pull_user_rows(self, st) -> None:
        """user_full[j, g*dg:(g+1)*dg] = final_g[row of my j-th user, :] for every rank g (gcf_peer_gather_cols on stream st)."""
        if self.ub == 0:
            return
        dg, G = self.dg, self.world
        if self.owner_major:       # my users are rows [rank * ubm, rank * ubm + ub) of every slice: contiguous remote reads
            src, ld_src = self._pb_final.pointers(self.rank * self.ubm * dg), dg
        else:                      # rows rank, rank + G, ...: pieces of dg floats at a pitch of G * dg
            src, ld_src = self._pb_final.pointers(self.rank * dg), G * dg
        _lib.check(self.lib.gcf_peer_gather_cols(src, G, ld_src, _lib.ptr(self.user_full), self.d_full, self.ub, dg, st),
                   "gcf_peer_gather_cols")

    def _push_user_grads(self, st) -> None:
        """owner-major only: g_final_g[rank * ubm + j, :] = g_user_full[j, lo_g:hi_g] for every rank g -- remote stores into ONE
        contiguous block per peer (call before the barrier that precedes the backward propagation)."""
        if self.ub == 0:
            return
        dg, G = self.dg, self.world
        src = _lib.ptr_values([self.g_user_full.data_ptr() + 4 * g * dg for g in range(G)])
        dst = _lib.ptr_values([self._pb_gfinal.base[g] + 4 * self.rank * self.ubm * dg for g in range(G)])
        _lib.check(self.lib.gcf_peer_copy2d(src, dst, self._rows_mine64, G, self.d_full, dg, dg, 0, st), "gcf_peer_copy2d")

    def _pull_user_grads(self, st) -> None:
        """natural numbering only: g_final[g + j G, :] = g_user_full_g[j, lo:hi] (call after the barrier that follows the BPR)."""
        dg, G = self.dg, self.world
        _lib.check(self.lib.gcf_peer_copy_blocks(self._pb_gu.pointers(self.lo), self._block_rows64, self._user_dst_off64, G,
                                                 self.d_full, _lib.ptr(self.g_final[:self.n_users]), G * dg, dg, st),
                   "gcf_peer_copy_blocks")

    phase_marks: Optional[list] = None   # tools/phase_dist.py: a list that receives (label, CUDA event) pairs of one step

    def _mark(self, label: str) -> None:
        if self.phase_marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.phase_marks.append((label, ev))

    def _loss_on_scores(self, neg_items: Optional[torch.Tensor]) -> torch.Tensor:
        """loss_layout="scores": partial scores on the local columns, one all-reduce of E floats, local gradient pass."""
        lib, st, dg, u = self.lib, _lib.current_stream(), self.dg, self.n_users
        if neg_items is None:
            _lib.check(lib.gcf_sample_negatives(self.seed, self.step_count, None, self.n_triples, 1, self.n_items, None, None, 1,
                                                _lib.ptr(self.neg), st), "gcf_sample_negatives")
            neg = self.neg
        else:
            neg = neg_items.to(torch.int64).reshape(-1)
            if self.order is not None:
                neg = neg[self.order]
            neg = neg.contiguous()
        ue, ie = self.final[:u], self.final[u:]
        args = (_lib.ptr(ue), dg, _lib.ptr(ie), dg, dg, _lib.ptr(self.pos_u), _lib.ptr(self.pos_i), _lib.ptr(neg), self.n_triples, 1)
        # 1. partial scores over the local columns (+ the local part of the squared-norm regulariser)
        _lib.check(lib.gcf_bpr_fwd(*args, _lib.BPR_RAW_SCORE, 0.0, _lib.REDUCE_SUM, self.reg, self.reg, 0.0, _lib.ptr(self.loss_reg),
                                   _lib.ptr(self.scores), _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st), "gcf_bpr_fwd")
        self._mark("sampler + partial scores")
        # 2. the one data-path collective of the step
        dist.all_reduce(self.scores, op=dist.ReduceOp.SUM)
        self._mark("score all-reduce")
        # 3. pointwise loss on the complete scores (redundantly on every rank: E floats)
        _lib.check(lib.gcf_bpr_coef_from_scores(_lib.ptr(self.scores), self.n_triples, _lib.BPR_SOFTPLUS, 0.0, _lib.REDUCE_MEAN,
                                                _lib.ptr(self.loss_pt), _lib.ptr(self.coef), _lib.ptr(self.bpr_ws),
                                                self.bpr_ws_bytes, st), "gcf_bpr_coef_from_scores")
        self._mark("coefficients")
        # 4. gradients w.r.t. the local columns
        self.g_final.zero_()
        _lib.check(lib.gcf_bpr_bwd(*args, _lib.ptr(self.coef), None, self.reg, self.reg, 0.0, _lib.ptr(self.g_final[:u]), dg,
                                   _lib.ptr(self.g_final[u:]), dg, st), "gcf_bpr_bwd")
        self._mark("gradient pass")
        return self.loss_pt / self.world + self.loss_reg   # the pointwise part is replicated: count it once over the ranks

    def _exchange_items(self):
        """item slices [I, d/G] -> [G, I, d/G] on every rank (asynchronous: queued on NCCL's stream)."""
        return dist.all_gather_into_tensor(self.item_blk.view(-1), self.final[self.n_users:].reshape(-1), async_op=True)

    def _exchange_users(self):
        """user slices [U, d/G] -> the [G, Ub, d/G] slices of this rank's user block (asynchronous)."""
        return dist.all_to_all_single(self.user_blk, self.final[:self.n_users], output_split_sizes=[self.ub] * self.world,
                                      input_split_sizes=self.block_rows, async_op=True)

    def _loss_on_rows_peer(self, neg_items: Optional[torch.Tensor]) -> torch.Tensor:
        """loss_layout="rows", exchange="peer": the same dataflow as _loss_on_rows, but the slices are pulled out of the peers'
        memory (csrc/peer.cu) instead of being sent by NCCL and re-laid out afterwards:

            barrier (every rank's slice of `final` is complete; nobody still reads last step's gradient partials)
            item_full[i, g*dg:(g+1)*dg] = final_g[U + i, :]         for all g    (gcf_peer_gather_cols over NVLink)
            user_full[j, g*dg:(g+1)*dg] = final_g[row of my j-th user, :]   my users: u % G == rank; with owner-major user rows
                                                                     (default) they are ONE contiguous block of every slice
            fused BPR on my E/G triples -> g_item_full [I, d], g_user_full [Ub, d]
            stage_g[rank][i, :] = g_item_full[i, lo_g:hi_g]          for all g    (gcf_peer_copy2d: remote stores)
            owner-major: g_final_g[my block + j, :] = g_user_full[j, lo_g:hi_g]   (gcf_peer_copy2d: remote stores)
            barrier (every rank's partials are complete and delivered)
            g_final[U + i, :] = sum_g stage[g][i, :]                 fixed order g = 0..G-1 (gcf_peer_sum_cols, local)
            node-order user rows: g_final[g + j G, :] = g_user_full_g[j, lo:hi]   (gcf_peer_copy_blocks: remote loads)

        Who may touch what when: a rank reads a peer's `final` only between barrier 1 and its own arrival at barrier 2 of the same
        step, and the peer rewrites `final` only after it has passed barrier 2; staging slices and the user rows of `g_final` are
        written by peers between barrier 1 and barrier 2 and read by their owner after barrier 2 and before it arrives at the next
        barrier 1.  Two barriers per step are therefore enough.
        """
        lib, st, dg, d, G, u, ub = self.lib, _lib.current_stream(), self.dg, self.d_full, self.world, self.n_users, self.ub
        if neg_items is None:
            # Philox slot = position in the user-major list of ALL triples: the draws do not depend on the number of ranks
            if self.n_local > 0:
                _lib.check(lib.gcf_sample_negatives_pos(self.seed, self.step_count, _lib.ptr(self.positions), self.n_local, 1,
                                                        self.n_items, _lib.ptr(self.neg), st), "gcf_sample_negatives_pos")
            neg = self.neg
        else:
            neg = neg_items.to(torch.int64).reshape(-1)
            if self.order is not None:
                neg = neg[self.order]
            neg = neg[self.positions].contiguous()
        self._mark("sampler")
        self._barrier()
        self._mark("barrier 1")
        # only now may the partials be cleared: before the barrier a slow peer could still be summing last step's
        self.g_item_full.zero_()
        self.g_user_full.zero_()
        self.loss_pt.zero_()
        self._mark("memsets")
        _lib.check(lib.gcf_peer_gather_cols(self._pb_final.pointers(u * dg), G, dg, _lib.ptr(self.item_full), d, self.n_items, dg, st),
                   "gcf_peer_gather_cols")
        self._mark("item slices -> rows (peer gather)")
        self._pull_user_rows(st)
        self._mark("user slices -> rows (peer gather)")
        w = 1.0 / max(self.n_triples, 1)
        if self.n_local > 0:
            _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(self.user_full), d, _lib.ptr(self.item_full), d, d, _lib.ptr(self.loc_u),
                                           _lib.ptr(self.loc_i), _lib.ptr(neg), self.n_local, 1, _lib.BPR_SOFTPLUS, 0.0,
                                           _lib.REDUCE_SUM, self.reg / w, self.reg / w, 0.0, w, _lib.ptr(self.loss_pt), None,
                                           _lib.ptr(self.g_user_full), d, _lib.ptr(self.g_item_full), d,
                                           _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st), "gcf_bpr_fwd_bwd")
        self._mark("fused BPR")
        # my partial's column piece g -> slot `rank` of rank g's staging buffer (remote stores, contiguous slices)
        src = _lib.ptr_values([self.g_item_full.data_ptr() + 4 * g * dg for g in range(G)])
        dst = _lib.ptr_values([self._pb_stage.base[g] + 4 * self.rank * self.n_items * dg for g in range(G)])
        _lib.check(lib.gcf_peer_copy2d(src, dst, self._rows_items64, G, d, dg, dg, 0, st), "gcf_peer_copy2d")
        self._mark("item gradients: column pieces -> owners (peer push)")
        if self.owner_major:
            self._push_user_grads(st)
            self._mark("user gradients: column pieces -> owners (peer push)")
        self._barrier()
        self._mark("barrier 2")
        stage = self._pb_stage.tensor
        _lib.check(lib.gcf_peer_sum_cols(_lib.ptr_values([stage.data_ptr() + 4 * g * self.n_items * dg for g in range(G)]), G, dg,
                                         _lib.ptr(self.g_final[u:]), dg, self.n_items, dg, st), "gcf_peer_sum_cols")
        self._mark("item gradients: sum of the G staged slices (local)")
        if not self.owner_major:
            self._pull_user_grads(st)
            self._mark("user gradients: rows -> my slice (peer copy)")
        return self.loss_pt * w

    def close(self) -> None:
        """Release the peer-visible buffers (collective: every rank, after its last step)."""
        if getattr(self, "_pb_final", None) is not None:
            torch.cuda.synchronize()
            dist.barrier()
            for pb in (self._pb_final, self._pb_stage, self._pb_gu, self._barrier, getattr(self, "_pb_gfinal", None)):
                if pb is not None:
                    pb.close()
            self._pb_final = self._pb_stage = self._pb_gu = None
            self.final = self.g_user_full = None

    def _loss_on_rows(self, neg_items: Optional[torch.Tensor], pending=None, defer_wait: bool = False):
        """loss_layout="rows": exchange column slices for full rows, fused BPR on this rank's user block, gradients back
        to column slices.  Fills self.g_final ([N, d/G]); returns this rank's share of the global loss (and, with
        defer_wait, the two pending gradient exchanges (users, items) instead of waiting for them)."""
        lib, st, dg, d, G, u, ub = self.lib, _lib.current_stream(), self.dg, self.d_full, self.world, self.n_users, self.ub
        if pending is not None:                 # overlap: the forward already queued both exchanges
            w_items, w_users = pending
        else:
            w_items, w_users = self._exchange_items(), self._exchange_users()
        if neg_items is None:
            # Philox slot = position in the user-major list of ALL triples: this rank draws the window [t_lo, t_hi) of the
            # stream the single-GPU trainer (and the "scores" layout) draws, whatever the number of ranks
            if self.n_local > 0:
                _lib.check(lib.gcf_sample_negatives_at(self.seed, self.step_count, self.t_lo, None, self.n_local, 1,
                                                       self.n_items, None, None, 1, _lib.ptr(self.neg), st), "gcf_sample_negatives_at")
            neg = self.neg
        else:
            neg = neg_items.to(torch.int64).reshape(-1)
            if self.order is not None:
                neg = neg[self.order]
            neg = neg[self.t_lo:self.t_hi].contiguous()
        self.g_item_full.zero_()
        self.g_user_full.zero_()
        self.loss_pt.zero_()
        self._mark("sampler + memsets")
        w_items.wait()
        self._mark("item all-gather (exposed)")
        _lib.check(lib.gcf_slices_to_rows(_lib.ptr(self.item_blk), _lib.ptr(self.item_full), d, self.n_items, G, dg, st),
                   "gcf_slices_to_rows")
        self._mark("item slices -> rows")
        w_users.wait()
        self._mark("user all-to-all (exposed)")
        if ub > 0:
            _lib.check(lib.gcf_slices_to_rows(_lib.ptr(self.user_blk), _lib.ptr(self.user_full), d, ub, G, dg, st),
                       "gcf_slices_to_rows")
        self._mark("user slices -> rows")
        w = 1.0 / max(self.n_triples, 1)
        if self.n_local > 0:
            # per-rank partial of the global mean: reduction = sum, loss and gradients scaled by 1/E (as in the row layout)
            _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(self.user_full), d, _lib.ptr(self.item_full), d, d, _lib.ptr(self.loc_u),
                                           _lib.ptr(self.loc_i), _lib.ptr(neg), self.n_local, 1, _lib.BPR_SOFTPLUS, 0.0,
                                           _lib.REDUCE_SUM, self.reg / w, self.reg / w, 0.0, w, _lib.ptr(self.loss_pt), None,
                                           _lib.ptr(self.g_user_full), d, _lib.ptr(self.g_item_full), d,
                                           _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st), "gcf_bpr_fwd_bwd")
        self._mark("fused BPR")
        # the (small) user exchange goes first: the item rows of the first backward product only need the user gradients
        if ub > 0:
            _lib.check(lib.gcf_rows_to_slices(_lib.ptr(self.g_user_full), d, _lib.ptr(self.user_blk), ub, G, dg, st),
                       "gcf_rows_to_slices")
        w_users = dist.all_to_all_single(self.g_final[:u], self.user_blk, output_split_sizes=self.block_rows,
                                         input_split_sizes=[ub] * G, async_op=True)
        _lib.check(lib.gcf_rows_to_slices(_lib.ptr(self.g_item_full), d, _lib.ptr(self.item_blk), self.n_items, G, dg, st),
                   "gcf_rows_to_slices")
        w_items = dist.reduce_scatter_tensor(self.g_final[u:].reshape(-1), self.item_blk.view(-1), op=dist.ReduceOp.SUM,
                                             async_op=True)
        self._mark("rows -> slices (users, items)")
        if defer_wait:
            return self.loss_pt * w, (w_users, w_items)
        w_users.wait()
        w_items.wait()
        self._mark("gradient all-to-all + reduce-scatter (exposed)")
        return self.loss_pt * w

    def index_buffers(self) -> List[torch.Tensor]:
        """The device tensors that hold this rank's training triples -- the per-step INPUT of the step (bench.py's e2e arm
        refills them from pinned host memory every step)."""
        return [self.loc_u, self.loc_i] if self.loss_layout == "rows" else [self.pos_u, self.pos_i]

    def step(self, neg_items: Optional[torch.Tensor] = None, marks: Optional[list] = None, wait_before_loss=None) -> torch.Tensor:
        """One optimisation step; returns the (global) loss.  neg_items: optional pre-drawn item ids for ALL triples in
        their original order (parity tests), identical on every rank.  wait_before_loss: optional CUDA event the loss phase
        waits for (an asynchronous refill of index_buffers())."""
        lib, st, g, dg, K, u = self.lib, _lib.current_stream(), self.graph, self.dg, self.k, self.n_users
        self.step_count += 1
        if marks is not None:
            e0, e1, e2, e3 = (torch.cuda.Event(enable_timing=True) for _ in range(4))
            e0.record()
        if self.overlap:
            run = self._step_overlapped_peer if self.exchange == "peer" else self._step_overlapped
            return run(neg_items, marks, (e0, e1, e2, e3) if marks is not None else None, wait_before_loss)
        self._mark("start")
        _lib.check(lib.gcf_propagate_fwd(g.struct_ref(), dg, K, _lib.ptr(self.table), _lib.ptr_array(self.layers),
                                         _lib.ptr(self.final), 1.0, _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_fwd")
        self._mark("forward propagation (K SpMM)")
        if marks is not None:
            e1.record()
        if wait_before_loss is not None:
            torch.cuda.current_stream().wait_event(wait_before_loss)
        if self.loss_layout == "rows":
            loss_local = self._loss_on_rows_peer(neg_items) if self.exchange == "peer" else self._loss_on_rows(neg_items)
        else:
            loss_local = self._loss_on_scores(neg_items)
        if marks is not None:
            e2.record()
        # the Adam update of the local [N, d/G] slice rides in the epilogue of the last backward SpMM
        _lib.check(lib.gcf_propagate_bwd_adam(g.struct_ref(), dg, K, _lib.ptr(self.g_final), None, 1.0, _lib.ptr(self.ping),
                                              _lib.ptr(self.pong), None, _lib.ptr(self.table), _lib.ptr(self.exp_avg),
                                              _lib.ptr(self.exp_avg_sq), self.lr, 0.9, 0.999, 1e-8, 0.0, 0, self.step_count,
                                              _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_bwd_adam")
        if marks is not None:
            e3.record()
            marks.append((e0, e1, K))
            marks.append((e2, e3, K))
        self._mark("backward propagation + Adam (K SpMM)")
        loss = loss_local.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)   # every rank holds 1/G of the global value ("scores") or its block's share ("rows")
        return loss

    def _block_spmm(self, graph: CSRGraph, ws, ws_bytes, x: torch.Tensor, out: torch.Tensor, addends, what: str) -> None:
        """out = A_block x + sum(addends): one launch on a row block of the operator."""
        lib, dg = self.lib, self.dg
        _lib.check(lib.gcf_spmm_csr_f32(graph.struct_ref(), dg, _lib.ptr(x), dg, None, 0, _lib.ptr(out), dg, _lib.EPILOGUE_NONE,
                                        1.0, 1.0, len(addends), _lib.ptr_array(list(addends)), _lib.float_array([1.0] * len(addends)),
                                        _lib.ptr(ws), ws_bytes, 0, _lib.current_stream()), what)

    def _step_overlapped(self, neg_items, marks, events, wait_before_loss=None) -> torch.Tensor:
        """The step with the exchanges of the loss hidden behind row blocks of the neighbouring propagation layers."""
        lib, st, g, dg, K, u = self.lib, _lib.current_stream(), self.graph, self.dg, self.k, self.n_users
        # ---- forward: K-1 full layers, the last one (with the layer sum) item rows first ----
        cur = self.table
        for k in range(K - 1):
            _lib.check(lib.gcf_spmm_csr_f32(g.struct_ref(), dg, _lib.ptr(cur), dg, _lib.ptr(self.layers[k]), dg, None, 0,
                                            _lib.EPILOGUE_NONE, 1.0, 1.0, 0, _lib.ptr_array([]), _lib.float_array([]),
                                            _lib.ptr(self.ws), self.ws_bytes, 0, st), "gcf_spmm_csr_f32")
            cur = self.layers[k]
        prev = [self.table] + self.layers[: K - 1]                     # E(0) .. E(K-1): final = sum of these + A E(K-1)
        self._block_spmm(self.g_items, self.ws_i, self.ws_i_bytes, cur, self.final[u:], [t[u:] for t in prev], "gcf_spmm_csr_f32")
        w_items = self._exchange_items()                                # on the wire while the user rows are computed
        self._block_spmm(self.g_users, self.ws_u, self.ws_u_bytes, cur, self.final[:u], [t[:u] for t in prev], "gcf_spmm_csr_f32")
        w_users = self._exchange_users()
        if events is not None:
            events[1].record()
        if wait_before_loss is not None:
            torch.cuda.current_stream().wait_event(wait_before_loss)
        loss_local, (g_users_done, g_items_done) = self._loss_on_rows(neg_items, pending=(w_items, w_users), defer_wait=True)
        if events is not None:
            events[2].record()
        # ---- backward: P = A g_final block by block, then the remaining K-1 layers with G(K-1) = g_final + P ----
        g_users_done.wait()                                             # item rows of P gather user gradients only
        self._block_spmm(self.g_items, self.ws_i, self.ws_i_bytes, self.g_final, self.p_buf[u:], [], "gcf_spmm_csr_f32")
        g_items_done.wait()                                             # the item reduce-scatter ran meanwhile
        self._block_spmm(self.g_users, self.ws_u, self.ws_u_bytes, self.g_final, self.p_buf[:u], [], "gcf_spmm_csr_f32")
        extra = [None] * (K - 1) + [self.p_buf]
        _lib.check(lib.gcf_propagate_bwd_adam(g.struct_ref(), dg, K - 1, _lib.ptr(self.g_final), _lib.ptr_array(extra), 1.0,
                                              _lib.ptr(self.ping), _lib.ptr(self.pong), None, _lib.ptr(self.table),
                                              _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.lr, 0.9, 0.999, 1e-8, 0.0, 0,
                                              self.step_count, _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_bwd_adam")
        if events is not None:
            events[3].record()
            marks.append((events[0], events[1], K))
            marks.append((events[2], events[3], K))
        loss = loss_local.clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        return loss

    def _step_overlapped_peer(self, neg_items, marks, events, wait_before_loss=None) -> torch.Tensor:
        """exchange="peer" with the item-side transfers hidden behind row blocks of the adjacent propagation layers.

        The operator is bipartite: item rows gather user columns only and vice versa.  So the LAST forward layer is launched
        as its item-row block first -- the moment every rank has its item slice (barrier on the side stream) the peer gather
        of the item table runs on the side stream while the main stream computes the user-row block -- and the FIRST backward
        product as its item-row block (needs the user gradients only) while the side stream pushes the item-gradient pieces
        to their owners and sums them.  The peer movers run ~128 CTAs and touch little local HBM, so they share the GPU with
        the SpMM at small cost.  Four stream-ordered barriers per step instead of two."""
        lib, g, dg, d, G, K, u, ub = self.lib, self.graph, self.dg, self.d_full, self.world, self.k, self.n_users, self.ub
        main, side = torch.cuda.current_stream(), self._side
        st = main.cuda_stream
        self._mark("start")
        # ---- forward: K-1 full layers, then the last one (with the layer sum) item rows first ----
        cur = self.table
        for k in range(K - 1):
            _lib.check(lib.gcf_spmm_csr_f32(g.struct_ref(), dg, _lib.ptr(cur), dg, _lib.ptr(self.layers[k]), dg, None, 0,
                                            _lib.EPILOGUE_NONE, 1.0, 1.0, 0, _lib.ptr_array([]), _lib.float_array([]),
                                            _lib.ptr(self.ws), self.ws_bytes, 0, st), "gcf_spmm_csr_f32")
            cur = self.layers[k]
        prev = [self.table] + self.layers[: K - 1]
        self._block_spmm(self.g_items, self.ws_i, self.ws_i_bytes, cur, self.final[u:], [t[u:] for t in prev], "gcf_spmm_csr_f32")
        items_ready = torch.cuda.Event()
        items_ready.record(main)
        with torch.cuda.stream(side):
            side.wait_event(items_ready)
            self._barrier()                      # every rank's item slice of `final` is complete
            _lib.check(lib.gcf_peer_gather_cols(self._pb_final.pointers(u * dg), G, dg, _lib.ptr(self.item_full), d, self.n_items, dg,
                                                side.cuda_stream), "gcf_peer_gather_cols")
            items_gathered = torch.cuda.Event()
            items_gathered.record(side)
        self._block_spmm(self.g_users, self.ws_u, self.ws_u_bytes, cur, self.final[:u], [t[:u] for t in prev], "gcf_spmm_csr_f32")
        self._mark("forward propagation (last layer in two row blocks; item gather on the side stream)")
        if events is not None:
            events[1].record()
        if wait_before_loss is not None:
            main.wait_event(wait_before_loss)
        # ---- loss ----
        if neg_items is None:
            if self.n_local > 0:
                _lib.check(lib.gcf_sample_negatives_pos(self.seed, self.step_count, _lib.ptr(self.positions), self.n_local, 1,
                                                        self.n_items, _lib.ptr(self.neg), st), "gcf_sample_negatives_pos")
            neg = self.neg
        else:
            neg = neg_items.to(torch.int64).reshape(-1)
            if self.order is not None:
                neg = neg[self.order]
            neg = neg[self.positions].contiguous()
        self._barrier()                          # every rank's user slice is complete; last step's partials are consumed
        self.g_item_full.zero_()
        self.g_user_full.zero_()
        self.loss_pt.zero_()
        self._pull_user_rows(st)
        main.wait_event(items_gathered)
        self._mark("sampler, barrier, memsets, user gather (+ wait for the item gather)")
        w = 1.0 / max(self.n_triples, 1)
        if self.n_local > 0:
            _lib.check(lib.gcf_bpr_fwd_bwd(_lib.ptr(self.user_full), d, _lib.ptr(self.item_full), d, d, _lib.ptr(self.loc_u),
                                           _lib.ptr(self.loc_i), _lib.ptr(neg), self.n_local, 1, _lib.BPR_SOFTPLUS, 0.0,
                                           _lib.REDUCE_SUM, self.reg / w, self.reg / w, 0.0, w, _lib.ptr(self.loss_pt), None,
                                           _lib.ptr(self.g_user_full), d, _lib.ptr(self.g_item_full), d,
                                           _lib.ptr(self.bpr_ws), self.bpr_ws_bytes, st), "gcf_bpr_fwd_bwd")
        self._mark("fused BPR")
        if events is not None:
            events[2].record()
        # ---- backward: P = A g_final block by block, the item gradients travel while the item rows of P are computed ----
        if self.owner_major:
            self._push_user_grads(st)                          # delivered into the owners' slices before the barrier
        self._barrier()                                        # every rank's BPR is done (and its user gradients delivered)
        if not self.owner_major:
            self._pull_user_grads(st)
        users_done = torch.cuda.Event()
        users_done.record(main)
        with torch.cuda.stream(side):
            # starts once the user copy has left the links, i.e. together with the item-row block below
            side.wait_event(users_done)
            src = _lib.ptr_values([self.g_item_full.data_ptr() + 4 * q * dg for q in range(G)])
            dst = _lib.ptr_values([self._pb_stage.base[q] + 4 * self.rank * self.n_items * dg for q in range(G)])
            _lib.check(lib.gcf_peer_copy2d(src, dst, self._rows_items64, G, d, dg, dg, 0, side.cuda_stream), "gcf_peer_copy2d")
            self._barrier()                                    # every rank's pieces have been delivered
            stage = self._pb_stage.tensor
            _lib.check(lib.gcf_peer_sum_cols(_lib.ptr_values([stage.data_ptr() + 4 * q * self.n_items * dg for q in range(G)]), G, dg,
                                             _lib.ptr(self.g_final[u:]), dg, self.n_items, dg, side.cuda_stream), "gcf_peer_sum_cols")
            item_grads = torch.cuda.Event()
            item_grads.record(side)
        self._mark("barrier + user gradients (peer copy)")
        self._block_spmm(self.g_items, self.ws_i, self.ws_i_bytes, self.g_final, self.p_buf[u:], [], "gcf_spmm_csr_f32")
        main.wait_event(item_grads)
        self._mark("item rows of A g (item gradients pushed + summed on the side stream)")
        self._block_spmm(self.g_users, self.ws_u, self.ws_u_bytes, self.g_final, self.p_buf[:u], [], "gcf_spmm_csr_f32")
        extra = [None] * (K - 1) + [self.p_buf]
        _lib.check(lib.gcf_propagate_bwd_adam(g.struct_ref(), dg, K - 1, _lib.ptr(self.g_final), _lib.ptr_array(extra), 1.0,
                                              _lib.ptr(self.ping), _lib.ptr(self.pong), None, _lib.ptr(self.table),
                                              _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.lr, 0.9, 0.999, 1e-8, 0.0, 0,
                                              self.step_count, _lib.ptr(self.ws), self.ws_bytes, st), "gcf_propagate_bwd_adam")
        self._mark("user rows of A g + remaining backward layers + Adam")
        if events is not None:
            events[3].record()
            marks.append((events[0], events[1], K))
            marks.append((events[2], events[3], K))
        loss = (self.loss_pt * w).clone()
        dist.all_reduce(loss, op=dist.ReduceOp.SUM)
        return loss

    def gathered_table(self) -> torch.Tensor:
        """[U + I, d] table in node order on every rank (for checks / evaluation)."""
        parts = [torch.empty_like(self.table) for _ in range(self.world)]
        dist.all_gather(parts, self.table)
        full = torch.cat(parts, dim=1)
        if not self.owner_major:
            return full
        users = full[self._user_row(torch.arange(self.true_users, device=full.device))]
        return torch.cat([users, full[self.n_users:]], dim=0)
