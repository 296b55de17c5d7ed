"""Text ingest on the GPU: the reference's data files -> device index arrays -> operators.

    read_text(path)                        file bytes -> uint8 CUDA tensor (pinned staging)
    parse_pairs(text, numeric=False)       first two tokens of every non-blank line as 8-byte keys     (load_data, ncl.py:542-543)
    DeviceInteraction.from_files(train, test, id_order="sorted" | "appearance" | "numeric")
        .user_num / .item_num / .users / .items (training pairs as dense indices, file order)          (Interaction._build)
        "sorted"     ids numbered by sorted() of the id strings          ncl.py:55-70, directau.py, mhcn.py, sept.py, buir.py
        "appearance" ids numbered by first appearance in the train file  selfcf.py:281-288, ssl4rec.py, diffnet.py
        "numeric"    ids are the integers in the file, counts = max + 1 over train and test            lightgcn.py:29-33
        .test_users / .test_items (dense indices, -1 for ids the training file does not contain)
        .norm_adj    raw bidirectional adjacency with duplicates kept (ncl.py:76-85) as a device CSR operator
        .interaction_mat, .sampler(), .user_ids() / .item_ids() (the id strings, decoded on demand)

Everything between the file bytes and the CSR operator runs in libgcf kernels (tokenising, key packing, radix sort + run
heads, binary-search lookup, the COO -> CSR build).  Ids of up to 8 bytes travel as one 64-bit key; longer ones (up to
MAX_ID_BYTES) as tuples of 64-bit words, sorted word by word (stable LSD radix passes) -- ncl.py:60-61 sorts arbitrary strings.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from . import _lib
from .graph import CSRGraph


MAX_ID_BYTES = 256      # 32 key words; the reference's datasets use ids of a few bytes


def _device(device=None) -> torch.device:
    return device if device is not None else torch.device("cuda", torch.cuda.current_device())


def read_text(path: str, device=None) -> torch.Tensor:
    """File content as a uint8 tensor on the GPU."""
    dev = _device(device)
    host = torch.from_numpy(np.fromfile(path, dtype=np.uint8))
    if host.numel() == 0:
        return torch.empty(0, dtype=torch.uint8, device=dev)
    return host.pin_memory().to(dev, non_blocking=True)


def parse_pairs(text: torch.Tensor, *, numeric: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """(first, second): int64 tensors holding the uint64 keys of the first two tokens of every record (see gcf.h);
    numeric=True parses decimal integers instead of packing the id bytes.  String ids longer than 8 bytes come back as
    [n_words, n_records] tensors (word-major key tuples, gcf_text_parse_pairs_words); otherwise the tensors are 1-D."""
    if not text.is_cuda or text.dtype != torch.uint8:
        raise RuntimeError("parse_pairs needs a uint8 CUDA tensor: recommendation_b200 has no CPU path")
    lib, st, dev = _lib.load(), _lib.current_stream(), text.device
    text = text.contiguous()
    n = int(text.numel())
    ws_bytes = lib.gcf_text_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    n_rec = torch.zeros(1, dtype=torch.int64, device=dev)
    _lib.check(lib.gcf_text_count_records(_lib.ptr(text) if n else None, n, _lib.ptr(n_rec), _lib.ptr(ws), ws_bytes, st),
               "gcf_text_count_records")
    r = int(n_rec.item())
    first = torch.empty(max(r, 1), dtype=torch.int64, device=dev)[:r]
    second = torch.empty(max(r, 1), dtype=torch.int64, device=dev)[:r]
    if numeric:
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        if r:
            _lib.check(lib.gcf_text_parse_pairs(_lib.ptr(text), n, 1, _lib.ptr(first), _lib.ptr(second), _lib.ptr(status),
                                                _lib.ptr(ws), ws_bytes, st), "gcf_text_parse_pairs")
        flags = int(status.item())
        if flags & 1:
            raise ValueError("parse_pairs: a line has fewer than two fields")
        if flags & 2:
            raise ValueError("parse_pairs: an id is not a decimal integer")
        return first, second
    status = torch.zeros(2, dtype=torch.int32, device=dev)
    if r:
        _lib.check(lib.gcf_text_parse_pairs_words(_lib.ptr(text), n, 1, r, _lib.ptr(first), _lib.ptr(second), _lib.ptr(status),
                                                  _lib.ptr(ws), ws_bytes, st), "gcf_text_parse_pairs_words")
    flags, longest = (int(v) for v in status.tolist())
    if flags & 1:
        raise ValueError("parse_pairs: a line has fewer than two fields")
    if flags & 2:                                        # an id longer than 8 bytes: parse again into key tuples
        if longest > MAX_ID_BYTES:
            raise ValueError(f"parse_pairs: an id of {longest} bytes exceeds MAX_ID_BYTES = {MAX_ID_BYTES}")
        w = -(-longest // 8)
        first = torch.empty(w, r, dtype=torch.int64, device=dev)
        second = torch.empty(w, r, dtype=torch.int64, device=dev)
        _lib.check(lib.gcf_text_parse_pairs_words(_lib.ptr(text), n, w, r, _lib.ptr(first), _lib.ptr(second), _lib.ptr(status),
                                                  _lib.ptr(ws), ws_bytes, st), "gcf_text_parse_pairs_words")
    return first, second


def _as_words(keys: torch.Tensor, n_words: int) -> torch.Tensor:
    """[n] or [w, n] keys -> [n_words, n] (zero words appended: shorter ids are zero padded anyway)."""
    k = keys.unsqueeze(0) if keys.dim() == 1 else keys
    if k.shape[0] < n_words:
        k = torch.cat([k, torch.zeros(n_words - k.shape[0], k.shape[1], dtype=k.dtype, device=k.device)], 0)
    return k.contiguous()


def sort_unique_words(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Key tuples [W, n] -> (distinct tuples ascending [W, m], first occurrence of each [m], dense index of every input [n])."""
    lib, st, dev = _lib.load(), _lib.current_stream(), keys.device
    w, n = keys.shape
    uniq = torch.empty(w, max(n, 1), dtype=torch.int64, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    rank = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    n_uniq = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib.gcf_sort_unique_words_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.gcf_sort_unique_words(_lib.ptr(keys.contiguous()) if n else None, w, n, _lib.ptr(uniq), _lib.ptr(first),
                                         _lib.ptr(n_uniq), _lib.ptr(rank), _lib.ptr(ws), ws_bytes, st), "gcf_sort_unique_words")
    m = int(n_uniq.item())
    return uniq[:, :m].contiguous(), first[:m].clone(), rank[:n].clone()


def lookup_words(table: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """Position of every key tuple ([W, n]) in the ascending `table` ([W, m]), -1 when absent."""
    lib = _lib.load()
    w, n = keys.shape
    m = int(table.shape[1])
    out = torch.empty(max(n, 1), dtype=torch.int64, device=keys.device)[:n]
    _lib.check(lib.gcf_lookup_sorted_words(_lib.ptr(table.contiguous()) if m else None, w, m, m, _lib.ptr(keys.contiguous()) if n else None,
                                           n, _lib.ptr(out) if n else None, _lib.current_stream()), "gcf_lookup_sorted_words")
    return out


def sort_unique(keys: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """(distinct keys ascending as unsigned 64-bit, position of each one's first occurrence)."""
    lib, st, dev = _lib.load(), _lib.current_stream(), keys.device
    n = int(keys.numel())
    uniq = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    first = torch.empty(max(n, 1), dtype=torch.int64, device=dev)
    n_uniq = torch.zeros(1, dtype=torch.int64, device=dev)
    ws_bytes = lib.gcf_sort_unique_workspace_bytes(n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(lib.gcf_sort_unique_u64(_lib.ptr(keys.contiguous()) if n else None, n, _lib.ptr(uniq), _lib.ptr(first),
                                       _lib.ptr(n_uniq), _lib.ptr(ws), ws_bytes, st), "gcf_sort_unique_u64")
    m = int(n_uniq.item())
    return uniq[:m].clone(), first[:m].clone()


def lookup(table: torch.Tensor, keys: torch.Tensor) -> torch.Tensor:
    """Position of every key in the ascending `table`, -1 when absent."""
    lib = _lib.load()
    out = torch.empty(max(keys.numel(), 1), dtype=torch.int64, device=keys.device)[: keys.numel()]
    _lib.check(lib.gcf_lookup_sorted_u64(_lib.ptr(table.contiguous()) if table.numel() else None, int(table.numel()),
                                         _lib.ptr(keys.contiguous()) if keys.numel() else None, int(keys.numel()),
                                         _lib.ptr(out) if keys.numel() else None, _lib.current_stream()), "gcf_lookup_sorted_u64")
    return out


def decode_keys(keys: torch.Tensor) -> List[str]:
    """The id strings of packed string keys, [n] or [W, n] (host side, for reports and `data.user`-style dictionaries)."""
    k = keys.unsqueeze(0) if keys.dim() == 1 else keys
    w, n = k.shape
    raw = np.ascontiguousarray(k.t().cpu().numpy().astype(">u8")).tobytes()       # record-major: 8 * w bytes per id
    return [raw[8 * w * j: 8 * w * (j + 1)].rstrip(b"\0").decode("ascii") for j in range(n)]


class _IdMap:
    """Dense numbering of one id column: `table` = distinct keys ascending, `rank[j]` = dense index of table[j].
    Keys are [n] (ids of up to 8 bytes) or [W, n] word tuples (longer ids)."""

    def __init__(self, keys: torch.Tensor, order: str):
        self.n_words = 1 if keys.dim() == 1 else int(keys.shape[0])
        if self.n_words == 1:
            self.table, first = sort_unique(keys.reshape(-1))
            n = int(self.table.numel())
        else:
            self.table, first, _ = sort_unique_words(keys)
            n = int(self.table.shape[1])
        if order == "appearance":                       # number the distinct ids by the position of their first occurrence
            by_first = torch.argsort(first)
            self.rank = torch.empty(n, dtype=torch.int64, device=keys.device)
            self.rank[by_first] = torch.arange(n, dtype=torch.int64, device=keys.device)
        else:
            self.rank = None                            # sorted(): the position in the table IS the index
        self.count = n

    def index(self, keys: torch.Tensor) -> torch.Tensor:
        kw = 1 if keys.dim() == 1 else int(keys.shape[0])
        if self.n_words == 1 and kw == 1:
            pos = lookup(self.table, keys.reshape(-1))
        else:
            w = max(self.n_words, kw)                   # e.g. a test file with longer ids than the training file
            pos = lookup_words(_as_words(self.table, w), _as_words(keys, w))
        if self.rank is None:
            return pos
        return torch.where(pos >= 0, self.rank[pos.clamp_min(0)], pos)

    def ids(self) -> List[str]:
        names = decode_keys(self.table)
        if self.rank is None:
            return names
        out = [""] * self.count
        for name, r in zip(names, self.rank.cpu().tolist()):
            out[r] = name
        return out


class DeviceInteraction:
    """The device-resident counterpart of the reference's `Interaction` objects (see the module docstring)."""

    def __init__(self, train_text: torch.Tensor, test_text: Optional[torch.Tensor] = None, *, id_order: str = "sorted"):
        if id_order not in ("sorted", "appearance", "numeric"):
            raise ValueError("id_order must be 'sorted', 'appearance' or 'numeric'")
        self.id_order = id_order
        numeric = id_order == "numeric"
        tu, ti = parse_pairs(train_text, numeric=numeric)
        eu, ei = parse_pairs(test_text, numeric=numeric) if test_text is not None else (tu[:0], ti[:0])
        if numeric:                                     # lightgcn.py:31-33: counts = max id over train and test, + 1
            self._umap = self._imap = None
            self.users, self.items, self.test_users, self.test_items = tu, ti, eu, ei
            mx = lambda a, b: int(max(int(a.max().item()) if a.numel() else -1, int(b.max().item()) if b.numel() else -1)) + 1
            self.user_num, self.item_num = mx(tu, eu), mx(ti, ei)
        else:
            self._umap, self._imap = _IdMap(tu, id_order), _IdMap(ti, id_order)
            self.user_num, self.item_num = self._umap.count, self._imap.count
            self.users, self.items = self._umap.index(tu), self._imap.index(ti)
            self.test_users, self.test_items = self._umap.index(eu), self._imap.index(ei)
        self.device = self.users.device
        self._norm_adj: Optional[CSRGraph] = None
        self._interaction: Optional[CSRGraph] = None
        self._gcf_sampler = None

    @classmethod
    def from_files(cls, train_path: str, test_path: Optional[str] = None, *, id_order: str = "sorted", device=None):
        return cls(read_text(train_path, device), read_text(test_path, device) if test_path else None, id_order=id_order)

    # ---- the reference's attributes -----------------------------------------------------
    @property
    def norm_adj(self) -> CSRGraph:
        """ncl.py:76-85: (u, i+U), (i+U, u) with value 1 per training line, duplicates kept (summed by the CSR build)."""
        if self._norm_adj is None:
            n = self.user_num + self.item_num
            rows = torch.cat([self.users, self.items + self.user_num])
            cols = torch.cat([self.items + self.user_num, self.users])
            self._norm_adj = CSRGraph.from_coo(rows, cols, None, n, n, norm="none", symmetric=True)
        return self._norm_adj

    def normalized_adj(self) -> CSRGraph:
        """selfcf.py:297-306 + 240-255: D^-1/2 (R + R^T) D^-1/2 with duplicate interactions summed."""
        return CSRGraph.from_pairs(self.users, self.items, self.user_num, self.item_num, norm="sym")

    @property
    def interaction_mat(self) -> CSRGraph:
        if self._interaction is None:
            self._interaction = CSRGraph.from_coo(self.users, self.items, None, self.user_num, self.item_num, norm="none")
        return self._interaction

    def edge_index(self) -> torch.Tensor:
        """lightgcn.py:36-39."""
        from .lightgcn import build_edge_index
        return build_edge_index(self.users, self.items, self.user_num)

    def sampler(self, seed: int = 0):
        from .sampling import PairwiseSampler
        if self._gcf_sampler is None:
            self._gcf_sampler = PairwiseSampler(self.users, self.items, self.user_num, self.item_num, seed=seed)
        return self._gcf_sampler

    def user_ids(self) -> List[str]:
        """id2user as a list (index -> id string); numeric mode: str(index)."""
        return [str(k) for k in range(self.user_num)] if self._umap is None else self._umap.ids()

    def item_ids(self) -> List[str]:
        return [str(k) for k in range(self.item_num)] if self._imap is None else self._imap.ids()
