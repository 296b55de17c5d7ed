"""Device-resident CSR adjacency and its construction through the integer kernels of libgcf.

Host-side mirror of the reference's graph containers:
  * `Interaction.__create_sparse_bipartite_adjacency` + `Graph.normalize_graph_mat`
    (selfcf.py:240-255,297-306; ssl4rec.py:79-88)                  -> CSRGraph.from_pairs(norm="sym")
  * `Interaction._build_adj` raw COO with duplicates (ncl.py:76-85, directau.py:132-141)
                                                                     -> CSRGraph.from_pairs(norm="none")
  * `load_data` edge_index + PyG gcn_norm (lightgcn.py:36-39,25)     -> CSRGraph.from_edge_index(norm="sym")
  * `TorchGraphInterface.convert_sparse_mat_to_tensor` (ncl.py:203-209, selfcf.py:219-225)
                                                                     -> CSRGraph.from_scipy
All index arithmetic (sort, duplicate merge, degrees) runs in CUDA kernels; only the long-row
schedule (a few thousand hub rows) is planned on the host from the row pointer.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib

_NORM_CODES = {"none": _lib.NORM_NONE, "sym": _lib.NORM_SYM, "row": _lib.NORM_ROW}
DEFAULT_CHUNK = int(os.environ.get("GCF_SPMM_CHUNK", "256"))


@dataclass
class LongRowPlan:
    """Rows longer than `chunk` entries, cut into chunks of `chunk` entries (host arrays, int32)."""

    chunk: int
    long_rows: np.ndarray       # [n_long]
    long_chunk_ptr: np.ndarray  # [n_long + 1]
    chunk_long: np.ndarray      # [n_chunks]

    @property
    def n_long(self) -> int:
        return int(self.long_rows.shape[0])

    @property
    def n_chunks(self) -> int:
        return int(self.chunk_long.shape[0])


def plan_long_rows(row_ptr: np.ndarray, chunk: int) -> LongRowPlan:
    """Degree-bucketed schedule: which rows are split, and into how many chunks (pure numpy)."""
    if chunk <= 0:
        raise ValueError("chunk must be positive")
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    deg = np.diff(row_ptr)
    long_rows = np.nonzero(deg > chunk)[0].astype(np.int32)
    n_chunks_per_row = -(-deg[long_rows] // chunk)
    long_chunk_ptr = np.zeros(long_rows.shape[0] + 1, dtype=np.int64)
    np.cumsum(n_chunks_per_row, out=long_chunk_ptr[1:])
    chunk_long = np.repeat(np.arange(long_rows.shape[0], dtype=np.int32), n_chunks_per_row)
    if long_chunk_ptr[-1] >= 2**31:
        raise ValueError("too many chunks")
    return LongRowPlan(chunk, long_rows, long_chunk_ptr.astype(np.int32), chunk_long.astype(np.int32))


def default_tile_nnz(nnz_short: int) -> int:
    """Entries per tile of the flat-stream SpMM: ~512 on large operators (the per-tile prologue -- tile -> row_ptr ->
    first (col, val) batch -- is amortised over 32 batches), smaller on small ones so that every SM still gets
    several tiles per resident sub-warp (148 SMs x 4 CTAs x 16 sub-warps x 2 waves)."""
    env = os.environ.get("GCF_SPMM_TILE")
    if env:
        return max(16, int(env))
    t = nnz_short // (148 * 4 * 16 * 2)
    return int(min(512, max(64, (t + 15) // 16 * 16)))


# hub columns per half of the operator; 0 = off, the default: measured on cfg1-cfg4 (profiles/r02_exp_hub_l1.log) the L1 residency
# hints are neutral at 4 CTAs/SM and 5-15 % slower at 3 CTAs/SM than the unflagged kernel, for every hub count tried
DEFAULT_HUBS = int(os.environ.get("GCF_SPMM_HUBS", "0"))
HUB_MAX_NNZ = int(os.environ.get("GCF_SPMM_HUB_MAX_NNZ", str(64_000_000)))
HUB_MIN_DEGREE = 128


def hub_flagged_columns(row_ptr: torch.Tensor, col_idx: torch.Tensor, n_cols: int, n_hubs: int,
                        min_degree: int = HUB_MIN_DEGREE) -> Optional[torch.Tensor]:
    """Copy of col_idx with bit 31 set on the entries that reference a hub column (gcf_csr_t.hub_col_idx).

    The rows are cut into two blocks holding half of the entries each (for the bipartite operators of the reference that
    is exactly user rows | item rows, whose entries reference item / user columns respectively); inside a block the hubs
    are its `n_hubs` most-referenced columns, provided they are referenced at least `min_degree` times (a row that each
    SM gathers less than about once is not worth keeping in its L1).  Build-time index arithmetic on the device (torch
    ops, once per operator); returns None when no column qualifies."""
    nnz = int(col_idx.numel())
    if nnz == 0 or n_hubs <= 0:
        return None
    half = int(torch.searchsorted(row_ptr.to(torch.int64), torch.tensor([nnz // 2], device=row_ptr.device), right=False)[0].item())
    cut = int(row_ptr[min(half, row_ptr.numel() - 1)].item())
    out = col_idx.clone()
    any_hub = False
    for e0, e1 in ((0, cut), (cut, nnz)):
        if e1 <= e0:
            continue
        cols = col_idx[e0:e1].to(torch.int64)
        deg = torch.bincount(cols, minlength=n_cols)
        top_deg, top_col = torch.topk(deg, min(n_hubs, n_cols))
        top_col = top_col[top_deg >= min_degree]
        if top_col.numel() == 0:
            continue
        is_hub = torch.zeros(n_cols, dtype=torch.bool, device=col_idx.device)
        is_hub[top_col] = True
        flag = is_hub[cols]
        out[e0:e1] = torch.where(flag, col_idx[e0:e1] | torch.tensor(-2**31, dtype=torch.int32, device=col_idx.device), col_idx[e0:e1])
        any_hub = True
    return out if any_hub else None


def plan_tiles(row_ptr: np.ndarray, chunk: int, tile_nnz: int):
    """Flat-stream schedule (pure numpy).  Returns (tiles, nz_rows, nz_row_ptr, empty_rows):

    * nz_rows int32 [n_nz]: ids of the rows with at least one entry (the COMPACT row numbering), nz_row_ptr int32
      [n_nz + 1] their row pointer, empty_rows int32 the other rows;
    * tiles int32 [n_tiles, 2] = (first, one-past-last) compact row of each tile.  Tiles are runs of consecutive
      compact rows of degree <= chunk (long rows belong to the chunk schedule and end a run); a run is cut wherever
      the running entry count of its rows crosses a multiple of tile_nnz, so a tile holds fewer than
      tile_nnz + chunk entries."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    deg_all = np.diff(row_ptr)
    nz_rows = np.nonzero(deg_all > 0)[0]
    empty_rows = np.nonzero(deg_all == 0)[0].astype(np.int32)
    nz_row_ptr = np.concatenate((row_ptr[nz_rows], row_ptr[-1:])) if nz_rows.size else np.zeros(1, np.int64)
    deg = deg_all[nz_rows]
    n = deg.shape[0]
    short = deg <= chunk
    if n == 0 or not short.any():
        return np.zeros((0, 2), dtype=np.int32), nz_rows.astype(np.int32), nz_row_ptr.astype(np.int32), empty_rows
    start_short = np.concatenate(([0], np.cumsum(np.where(short, deg, 0))[:-1]))   # short entries before each row
    bucket = start_short // max(int(tile_nnz), 1)
    prev_short = np.concatenate(([False], short[:-1]))
    prev_bucket = np.concatenate(([-1], bucket[:-1]))
    is_start = short & (~prev_short | (bucket != prev_bucket))
    starts = np.nonzero(is_start)[0]
    # a tile ends at the next tile start or at the next long row, whichever comes first
    stops = np.concatenate((np.nonzero(is_start | ~short)[0], [n]))
    ends = stops[np.searchsorted(stops, starts, side="right")]
    tiles = np.stack((starts, ends), axis=1).astype(np.int32)
    return tiles, nz_rows.astype(np.int32), nz_row_ptr.astype(np.int32), empty_rows


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: recommendation_b200 has no CPU path")


class CSRGraph:
    """CSR operator on the GPU (int32 structure, fp32 values) + long-row schedule + SpMM scratch."""

    def __init__(self, row_ptr: torch.Tensor, col_idx: torch.Tensor, vals: torch.Tensor, n_rows: int, n_cols: int,
                 *, symmetric: bool = False, chunk: Optional[int] = None, rowsum: Optional[torch.Tensor] = None,
                 dinv: Optional[torch.Tensor] = None, tile_nnz: Optional[int] = None, hubs: Optional[int] = None):
        for name, t, dt in (("row_ptr", row_ptr, torch.int32), ("col_idx", col_idx, torch.int32), ("vals", vals, torch.float32)):
            _require_cuda(t, name)
            if t.dtype != dt or not t.is_contiguous():
                raise ValueError(f"{name} must be contiguous {dt}")
        if row_ptr.numel() != n_rows + 1:
            raise ValueError("row_ptr must have n_rows + 1 entries")
        self.row_ptr, self.col_idx, self.vals = row_ptr, col_idx, vals
        self.n_rows, self.n_cols = int(n_rows), int(n_cols)
        self.nnz = int(col_idx.numel())
        self.symmetric = bool(symmetric)
        self.rowsum, self.dinv = rowsum, dinv
        self.device = row_ptr.device
        self._transpose: Optional["CSRGraph"] = None
        self._workspaces: Dict[int, torch.Tensor] = {}
        self.chunk = DEFAULT_CHUNK if chunk is None else int(chunk)
        row_ptr_host = row_ptr.cpu().numpy()
        plan = plan_long_rows(row_ptr_host, self.chunk)
        self.plan = plan
        n_short = self.nnz - int(np.diff(row_ptr_host.astype(np.int64))[plan.long_rows].sum()) if plan.n_long else self.nnz
        self.tile_nnz = default_tile_nnz(n_short) if tile_nnz is None else int(tile_nnz)
        if self.tile_nnz > 0 and self.nnz > 0:
            tiles, nz_rows, nz_row_ptr, empty = plan_tiles(row_ptr_host, self.chunk, self.tile_nnz)
        else:
            tiles, nz_rows, nz_row_ptr, empty = np.zeros((0, 2), np.int32), np.zeros(0, np.int32), np.zeros(1, np.int32), np.zeros(0, np.int32)
        self.n_tiles = int(tiles.shape[0])
        self.n_empty = int(empty.shape[0]) if self.n_tiles else 0
        self._tiles = torch.from_numpy(np.ascontiguousarray(tiles)).to(self.device)
        self._empty_rows = torch.from_numpy(empty).to(self.device) if self.n_empty else None
        # the compact numbering is the plain one when no row is empty
        self._nz_rows = torch.from_numpy(nz_rows).to(self.device) if self.n_empty else None
        self._nz_row_ptr = torch.from_numpy(nz_row_ptr).to(self.device) if self.n_empty else None
        self._long_rows = torch.from_numpy(plan.long_rows).to(self.device)
        self._long_chunk_ptr = torch.from_numpy(plan.long_chunk_ptr).to(self.device)
        self._chunk_long = torch.from_numpy(plan.chunk_long).to(self.device)
        # hub flags for the flat-stream kernel: operators small enough for their working set to live in the L2
        n_hubs = (DEFAULT_HUBS if self.nnz <= HUB_MAX_NNZ else 0) if hubs is None else int(hubs)
        self._hub_col_idx = hub_flagged_columns(row_ptr, col_idx, self.n_cols, n_hubs) if (n_hubs > 0 and self.n_tiles) else None
        self._struct = _lib.CsrStruct(
            n_rows=self.n_rows, n_cols=self.n_cols, nnz=self.nnz,
            row_ptr=row_ptr.data_ptr(), col_idx=col_idx.data_ptr() if self.nnz else None,
            vals=vals.data_ptr() if self.nnz else None,
            chunk=plan.chunk if plan.n_long else 0, n_long=plan.n_long, n_chunks=plan.n_chunks,
            long_rows=self._long_rows.data_ptr() if plan.n_long else None,
            long_chunk_ptr=self._long_chunk_ptr.data_ptr() if plan.n_long else None,
            chunk_long=self._chunk_long.data_ptr() if plan.n_long else None,
            tiles=self._tiles.data_ptr() if self.n_tiles else None, n_tiles=self.n_tiles, n_empty=self.n_empty,
            empty_rows=_lib.ptr(self._empty_rows), nz_row_ptr=_lib.ptr(self._nz_row_ptr), nz_rows=_lib.ptr(self._nz_rows),
            hub_col_idx=_lib.ptr(self._hub_col_idx),
        )

    # ---- construction ---------------------------------------------------------------------
    @classmethod
    def from_coo(cls, rows: torch.Tensor, cols: torch.Tensor, vals: Optional[torch.Tensor], n_rows: int, n_cols: int,
                 *, norm: str = "none", symmetric: bool = False, chunk: Optional[int] = None,
                 tile_nnz: Optional[int] = None, hubs: Optional[int] = None) -> "CSRGraph":
        """COO (int64 indices, duplicates allowed) -> canonical CSR, then value normalisation."""
        if norm not in _NORM_CODES:
            raise ValueError(f"norm must be one of {sorted(_NORM_CODES)}")
        lib = _lib.load()
        _require_cuda(rows, "rows")
        _require_cuda(cols, "cols")
        rows = rows.to(torch.int64).contiguous()
        cols = cols.to(torch.int64).contiguous()
        nnz_in = int(rows.numel())
        if cols.numel() != nnz_in:
            raise ValueError("rows and cols must have the same length")
        if vals is not None:
            _require_cuda(vals, "vals")
            vals = vals.to(torch.float32).contiguous()
            if vals.numel() != nnz_in:
                raise ValueError("vals must match rows/cols")
        dev = rows.device
        stream = _lib.current_stream()
        row_ptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        col_idx = torch.empty(max(nnz_in, 1), dtype=torch.int32, device=dev)
        out_vals = torch.empty(max(nnz_in, 1), dtype=torch.float32, device=dev)
        nnz_out = torch.zeros(1, dtype=torch.int64, device=dev)
        ws_bytes = lib.gcf_coo_to_csr_workspace_bytes(nnz_in, n_rows, n_cols)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gcf_coo_to_csr_stable(_lib.ptr(rows), _lib.ptr(cols), _lib.ptr(vals), nnz_in, n_rows, n_cols,
                                             _lib.ptr(row_ptr), _lib.ptr(col_idx), _lib.ptr(out_vals), _lib.ptr(nnz_out),
                                             _lib.ptr(ws), ws_bytes, stream), "gcf_coo_to_csr_stable")
        nnz = int(nnz_out.item())
        del ws
        if nnz < 0:
            raise ValueError(f"COO indices out of range for a {n_rows} x {n_cols} matrix (scipy / torch raise here too)")
        if 0 < nnz < col_idx.numel():  # duplicates were merged: release the slack
            col_idx, out_vals = col_idx[:nnz].clone(), out_vals[:nnz].clone()
        elif nnz == 0:                 # keep a non-null base pointer for empty operators
            col_idx, out_vals = col_idx[:0], out_vals[:0]
        rowsum = torch.empty(n_rows, dtype=torch.float32, device=dev)
        dinv = torch.empty(n_rows, dtype=torch.float32, device=dev)
        normed = torch.empty_like(out_vals)
        if n_rows > 0:
            _lib.check(lib.gcf_norm_values(_NORM_CODES[norm], _lib.ptr(row_ptr), _lib.ptr(col_idx), _lib.ptr(out_vals),
                                           n_rows, n_cols, _lib.ptr(normed), _lib.ptr(rowsum), _lib.ptr(dinv), stream),
                       "gcf_norm_values")
        return cls(row_ptr, col_idx, normed, n_rows, n_cols, symmetric=symmetric, chunk=chunk, rowsum=rowsum, dinv=dinv,
                   tile_nnz=tile_nnz, hubs=hubs)

    @classmethod
    def from_edge_index(cls, edge_index: torch.Tensor, num_nodes: int, *, norm: str = "sym",
                        symmetric: bool = True, chunk: Optional[int] = None, tile_nnz: Optional[int] = None,
                        hubs: Optional[int] = None) -> "CSRGraph":
        """edge_index [2, E'] (PyG convention: message flows row -> col, out[col] += w * x[row]).

        The operator applied to X is therefore M with M[col, row] = w, i.e. CSR rows = edge_index[1].
        For the bidirectional edge list of lightgcn.py:36-39 M is symmetric.
        """
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError("edge_index must have shape [2, E]")
        return cls.from_coo(edge_index[1], edge_index[0], None, num_nodes, num_nodes, norm=norm,
                            symmetric=symmetric, chunk=chunk, tile_nnz=tile_nnz, hubs=hubs)

    @classmethod
    def from_pairs(cls, users: torch.Tensor, items: torch.Tensor, n_users: int, n_items: int, *, norm: str = "sym",
                   chunk: Optional[int] = None, tile_nnz: Optional[int] = None, hubs: Optional[int] = None) -> "CSRGraph":
        """Bipartite user-item pairs -> symmetric (U+I)x(U+I) adjacency [[0,R],[R^T,0]]."""
        lib = _lib.load()
        _require_cuda(users, "users")
        users = users.to(torch.int64).contiguous()
        items = items.to(torch.int64).contiguous()
        e = int(users.numel())
        rows = torch.empty(2 * e, dtype=torch.int64, device=users.device)
        cols = torch.empty(2 * e, dtype=torch.int64, device=users.device)
        _lib.check(lib.gcf_bipartite_edge_index(_lib.ptr(users), _lib.ptr(items), e, n_users, _lib.ptr(rows), _lib.ptr(cols),
                                                _lib.current_stream()), "gcf_bipartite_edge_index")
        n = n_users + n_items
        return cls.from_coo(rows, cols, None, n, n, norm=norm, symmetric=True, chunk=chunk, tile_nnz=tile_nnz, hubs=hubs)

    @classmethod
    def from_scipy(cls, mat, *, norm: str = "none", device: Optional[torch.device] = None,
                   symmetric: Optional[bool] = None, chunk: Optional[int] = None, tile_nnz: Optional[int] = None,
                   hubs: Optional[int] = None) -> "CSRGraph":
        """Any scipy.sparse matrix (the reference's `data.norm_adj`); COO triplets are uploaded as they are
        (duplicates kept, like convert_sparse_mat_to_tensor) and canonicalised on the GPU."""
        coo = mat.tocoo()
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        rows = torch.from_numpy(np.ascontiguousarray(coo.row, dtype=np.int64)).to(dev)
        cols = torch.from_numpy(np.ascontiguousarray(coo.col, dtype=np.int64)).to(dev)
        vals = torch.from_numpy(np.ascontiguousarray(coo.data, dtype=np.float32)).to(dev)
        n_rows, n_cols = coo.shape
        if symmetric is None:
            symmetric = False
            if n_rows == n_cols:
                diff = (mat - mat.T)
                symmetric = diff.nnz == 0 or float(abs(diff).max()) == 0.0
        return cls.from_coo(rows, cols, vals, n_rows, n_cols, norm=norm, symmetric=symmetric, chunk=chunk, tile_nnz=tile_nnz,
                            hubs=hubs)

    # ---- derived operators ----------------------------------------------------------------
    def transpose(self) -> "CSRGraph":
        """CSR of the transpose (self when the operator is symmetric); cached."""
        if self.symmetric:
            return self
        if self._transpose is None:
            lib = _lib.load()
            dev = self.device
            t_row_ptr = torch.empty(self.n_cols + 1, dtype=torch.int32, device=dev)
            t_col_idx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[: self.nnz]
            t_vals = torch.empty(max(self.nnz, 1), dtype=torch.float32, device=dev)[: self.nnz]
            ws_bytes = lib.gcf_csr_transpose_workspace_bytes(self.nnz, self.n_rows, self.n_cols)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            _lib.check(lib.gcf_csr_transpose(_lib.ptr(self.row_ptr), _lib.ptr(self.col_idx), _lib.ptr(self.vals),
                                             self.n_rows, self.n_cols, self.nnz, _lib.ptr(t_row_ptr), _lib.ptr(t_col_idx),
                                             _lib.ptr(t_vals), _lib.ptr(ws), ws_bytes, _lib.current_stream()),
                       "gcf_csr_transpose")
            self._transpose = CSRGraph(t_row_ptr, t_col_idx, t_vals, self.n_cols, self.n_rows, chunk=self.chunk)
            self._transpose._transpose = self
        return self._transpose

    def with_values(self, vals: torch.Tensor, *, symmetric: bool = False) -> "CSRGraph":
        """Same sparsity pattern and long-row schedule, different values (shares row_ptr / col_idx / workspaces)."""
        if vals.shape != self.vals.shape or vals.dtype != torch.float32 or not vals.is_cuda:
            raise ValueError("with_values: need a float32 CUDA tensor with one value per stored entry")
        g = object.__new__(CSRGraph)
        g.__dict__.update(self.__dict__)
        g.__dict__.pop("_tperm", None)
        g.vals = vals.contiguous()
        g.symmetric = symmetric
        g._transpose = None
        g.rowsum = g.dinv = None
        g._struct = _lib.CsrStruct.from_buffer_copy(self._struct)
        g._struct.vals = g.vals.data_ptr() if self.nnz else None
        return g

    def transpose_permutation(self) -> torch.Tensor:
        """perm int32 [nnz]: entry j of the transposed CSR is entry perm[j] of this one (cached).  The stable transpose
        kernel moves the value payload bit for bit, so the entry ids ride through it re-interpreted as float32."""
        p = getattr(self, "_tperm", None)
        if p is None:
            ids = torch.arange(self.nnz, dtype=torch.int32, device=self.device).view(torch.float32)
            p = self.with_values(ids).transpose().vals.view(torch.int32)   # with_values() drops the "symmetric" shortcut
            self._tperm = p
        return p

    def dropout(self, rate: float, *, seed: int, offset: int = 0) -> "CSRGraph":
        """Entry-wise dropout of the operator (sparse_dropout, buir.py:300-309): every stored entry kept independently with
        probability 1 - rate and rescaled by 1 / (1 - rate).  The result carries its exact transpose (same mask), so
        autograd through functional.spmm / propagate is consistent.  The pattern is kept; dropped entries are zeros."""
        lib = _lib.load()
        st = _lib.current_stream()
        seed, offset = int(seed) & (2**64 - 1), int(offset) & (2**64 - 1)
        out = torch.empty_like(self.vals)
        _lib.check(lib.gcf_csr_dropout_values(_lib.ptr(self.vals), self.nnz, None, float(rate), seed, offset, _lib.ptr(out), st),
                   "gcf_csr_dropout_values")
        fwd = self.with_values(out)
        # the transpose keeps ITS pattern too (self for a symmetric operator); only the values move, through the permutation
        base_t, perm = self.transpose(), self.transpose_permutation()
        out_t = torch.empty_like(self.vals)
        _lib.check(lib.gcf_csr_dropout_values(_lib.ptr(self.vals), self.nnz, _lib.ptr(perm), float(rate), seed, offset,
                                              _lib.ptr(out_t), st), "gcf_csr_dropout_values")
        bwd = base_t.with_values(out_t)
        fwd._transpose, bwd._transpose = bwd, fwd
        return fwd

    # ---- plumbing for the kernels ---------------------------------------------------------
    @property
    def struct(self) -> "_lib.CsrStruct":
        return self._struct

    def struct_ref(self):
        return ctypes.byref(self._struct)

    def workspace(self, d: int) -> Tuple[Optional[torch.Tensor], int]:
        """Zero-initialised SpMM scratch for embedding width d (partial sums + self-resetting tickets)."""
        ws = self._workspaces.get(d)
        if ws is None:
            nbytes = _lib.load().gcf_spmm_workspace_bytes(self.struct_ref(), d)
            ws = torch.zeros(max(nbytes, 16), dtype=torch.uint8, device=self.device)
            self._workspaces[d] = ws
        return ws, ws.numel()

    def degrees(self) -> torch.Tensor:
        """Row sums of the un-normalised adjacency (integer-valued for 0/1 graphs)."""
        if self.rowsum is None:
            raise RuntimeError("this CSRGraph was not built through from_coo; no row sums recorded")
        return self.rowsum

    def to_scipy(self):
        import scipy.sparse as sp

        return sp.csr_matrix((self.vals.cpu().numpy(), self.col_idx.cpu().numpy(), self.row_ptr.cpu().numpy()),
                             shape=(self.n_rows, self.n_cols))

    def __repr__(self) -> str:
        return (f"CSRGraph({self.n_rows}x{self.n_cols}, nnz={self.nnz}, symmetric={self.symmetric}, "
                f"long_rows={self.plan.n_long}, chunks={self.plan.n_chunks})")


# =========================================================================================
# Reference-named helpers (selfcf.py:219-255 = ncl.py:30-44,203-209 = mhcn.py:47-84)
# =========================================================================================
class Graph:
    """`Graph.normalize_graph_mat(adj_mat)` with the reference's semantics -- square: D^-1/2 A D^-1/2, otherwise D^-1 A,
    inf -> 0 -- computed by the integer / normalisation kernels on the GPU.  Returns a scipy CSR matrix like the
    reference (callers store it as `data.norm_adj`); `normalize_to_device` skips the download."""

    @staticmethod
    def normalize_to_device(adj_mat, device: Optional[torch.device] = None) -> CSRGraph:
        shape = adj_mat.get_shape() if hasattr(adj_mat, "get_shape") else adj_mat.shape
        norm = "sym" if shape[0] == shape[1] else "row"
        return CSRGraph.from_scipy(adj_mat, norm=norm, device=device)

    @staticmethod
    def normalize_graph_mat(adj_mat):
        return Graph.normalize_to_device(adj_mat).to_scipy()


class TorchGraphInterface:
    """`convert_sparse_mat_to_tensor(X)`: the reference builds an uncoalesced torch.sparse COO tensor; here the result
    is the device CSR operator the SpMM kernels consume (use it with functional.spmm / propagate)."""

    @staticmethod
    def convert_sparse_mat_to_tensor(X, device: Optional[torch.device] = None) -> CSRGraph:
        return CSRGraph.from_scipy(X, norm="none", device=device)
