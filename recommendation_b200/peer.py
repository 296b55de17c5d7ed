"""Peer-visible device buffers for the multi-GPU trainers (one process per GPU on one NVSwitch box).

A `PeerBuffer` is a cudaMalloc block of this rank that every other rank of the process group has mapped into its own
address space (CUDA IPC, `gcf_peer_export` / `gcf_peer_open`), so that a kernel of rank a can read rank b's rows directly
over NVLink -- csrc/peer.cu's movers take the list of per-rank base pointers this class hands out.  The reference has no
distributed code; this replaces NCCL + layout-pass pairs of the loss exchange (dist.py).

Nothing here synchronises ranks: `stream_barrier()` is the ordering primitive the trainers use between a producer kernel on
one rank and a consumer kernel on another (a one-element all-reduce enqueued on the current stream).
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence

import torch
import torch.distributed as dist

from . import _lib

HANDLE_BYTES = 64


class _Block:
    """Owner of one gcf_peer_alloc block, exposed to torch through the CUDA array interface."""

    def __init__(self, n_floats: int):
        lib = _lib.load()
        out = ctypes.c_void_p()
        _lib.check(lib.gcf_peer_alloc(max(int(n_floats), 4) * 4, ctypes.byref(out)), "gcf_peer_alloc")
        self.ptr = int(out.value)
        self.__cuda_array_interface__ = {"shape": (max(int(n_floats), 4),), "typestr": "<f4", "data": (self.ptr, False),
                                         "version": 2, "strides": None}

    def free(self):
        if self.ptr:
            _lib.load().gcf_peer_free(ctypes.c_void_p(self.ptr))
            self.ptr = 0


class PeerBuffer:
    """fp32 [rows, cols] buffer of this rank + the addresses of every rank's buffer of the same role.

    Shapes may differ between ranks (user blocks have different sizes); `base[g]` is the device address of rank g's buffer as
    seen from THIS process (own allocation for g == rank, IPC mapping otherwise)."""

    def __init__(self, rows: int, cols: int, device: torch.device, group=None):
        if not dist.is_initialized():
            raise RuntimeError("PeerBuffer needs an initialised torch.distributed process group")
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.rows, self.cols = int(rows), int(cols)
        lib = _lib.load()
        self._block = _Block(self.rows * self.cols)
        flat = torch.as_tensor(self._block, device=device)
        if flat.data_ptr() != self._block.ptr:
            raise RuntimeError("PeerBuffer: torch copied the peer block instead of aliasing it")
        self.tensor = flat[: self.rows * self.cols].view(self.rows, self.cols)
        handle = ctypes.create_string_buffer(HANDLE_BYTES)
        _lib.check(lib.gcf_peer_export(ctypes.c_void_p(self._block.ptr), handle), "gcf_peer_export")
        handles: List[bytes] = [b""] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        self.base: List[int] = []
        self._opened: List[int] = []
        for g in range(self.world):
            if g == self.rank:
                self.base.append(self._block.ptr)
                continue
            out = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(handles[g], HANDLE_BYTES)
            _lib.check(lib.gcf_peer_open(buf, ctypes.byref(out)), "gcf_peer_open")
            self.base.append(int(out.value))
            self._opened.append(int(out.value))

    def pointers(self, float_offset: int = 0) -> "ctypes.Array":
        """Host array of the G base pointers, each advanced by `float_offset` floats (a row / column offset common to all ranks)."""
        arr = (ctypes.c_void_p * self.world)()
        for g, b in enumerate(self.base):
            arr[g] = b + 4 * int(float_offset)
        return arr

    def close(self) -> None:
        """Unmap the peers' blocks and free ours.  Collective in effect: call it on every rank, after a barrier."""
        lib = _lib.load()
        for p in self._opened:
            lib.gcf_peer_close(ctypes.c_void_p(p))
        self._opened = []
        self.tensor = None
        self._block.free()


_flag = {}


def stream_barrier(device: torch.device, group=None) -> None:
    """Every rank's work enqueued on its current stream BEFORE this call completes before any rank's work enqueued AFTER it
    starts: a one-element NCCL all-reduce (its kernel waits for the current stream and the current stream waits for it)."""
    key = (device.index, id(group))
    t = _flag.get(key)
    if t is None:
        t = _flag[key] = torch.zeros(1, dtype=torch.float32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)


def int64_array(values: Sequence[int]) -> "ctypes.Array":
    arr = (ctypes.c_int64 * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
