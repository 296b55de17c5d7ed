"""Peer-visible device buffers for the multi-GPU trainers (one process per GPU on one NVSwitch box).

A `PeerBuffer` is a cudaMalloc block of this rank that every other rank of the process group has mapped into its own
address space (CUDA IPC, `gcf_peer_export` / `gcf_peer_open`), so that a kernel of rank a can read rank b's rows directly
over NVLink -- csrc/peer.cu's movers take the list of per-rank base pointers this class hands out.  The reference has no
distributed code; this replaces NCCL + layout-pass pairs of the loss exchange (dist.py).

`PeerBarrier` is the ordering primitive the trainers use between a producer kernel on one rank and a consumer kernel on
another: a one-warp kernel over peer-visible flag words, enqueued on the current stream.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

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


class PeerBarrier:
    """Stream-ordered barrier of the ranks on the device (gcf_peer_barrier): every rank's work enqueued on its current stream
    BEFORE the call completes, and is visible to the peers, before any rank's work enqueued AFTER it starts.  A one-warp
    kernel over peer-visible flag words -- no NCCL kernel, so it also works from a high-priority side stream while the default
    stream is busy.  All ranks must call it in the same order (each call advances the epoch)."""

    def __init__(self, device: torch.device, group=None):
        self._flags = PeerBuffer(1, 64, device, group)        # 64 zero-initialised 32-bit words per rank
        self.rank, self.world = self._flags.rank, self._flags.world
        self._ptrs = self._flags.pointers()
        self._epoch = 0
        self._lib = _lib.load()

    def __call__(self, stream: Optional[int] = None) -> None:
        self._epoch += 1
        _lib.check(self._lib.gcf_peer_barrier(self._ptrs, self.world, self.rank, self._epoch & 0xFFFFFFFF,
                                              _lib.current_stream() if stream is None else stream), "gcf_peer_barrier")

    def close(self) -> None:
        self._flags.close()


_probe_result = {}


def available(device: torch.device, group=None) -> bool:
    """Collective probe: can every rank of the group map every other rank's memory (CUDA IPC + peer access)?  Every rank
    exports a small block, tries to open everybody else's and the ranks agree on the outcome (all-reduce of a flag), so that
    a box without peer access makes ALL ranks choose the NCCL exchange instead of leaving some of them stuck in a collective."""
    key = (device.index, id(group))
    if key in _probe_result:
        return _probe_result[key]
    import os

    if os.environ.get("GCF_PEER_DISABLE", "") not in ("", "0"):     # operator switch (set it for every rank)
        _probe_result[key] = False
        return False
    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    ok, why = True, ""
    block, payload = None, None
    try:
        block = _Block(64)
        handle = ctypes.create_string_buffer(HANDLE_BYTES)
        _lib.check(lib.gcf_peer_export(ctypes.c_void_p(block.ptr), handle), "gcf_peer_export")
        payload = handle.raw
    except Exception as e:   # noqa: BLE001 -- any failure means "not available"
        ok, why = False, str(e)
    handles: List[Optional[bytes]] = [None] * world
    dist.all_gather_object(handles, payload, group=group)
    opened = []
    if ok:
        for g in range(world):
            if g == rank:
                continue
            if handles[g] is None:
                ok, why = False, f"rank {g} could not export a block"
                break
            out = ctypes.c_void_p()
            buf = ctypes.create_string_buffer(handles[g], HANDLE_BYTES)
            if lib.gcf_peer_open(buf, ctypes.byref(out)) != 0:
                ok, why = False, _lib.last_error()
                break
            opened.append(int(out.value))
    flag = torch.tensor([1.0 if ok else 0.0], device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    torch.cuda.synchronize(device)
    for p in opened:
        lib.gcf_peer_close(ctypes.c_void_p(p))
    dist.barrier(group=group)            # nobody frees its block while a peer still has it mapped
    if block is not None:
        block.free()
    result = bool(flag.item() > 0.5)
    if not result and rank == 0:
        import sys

        print(f"[recommendation_b200.peer] peer memory is not available on this box ({why or 'another rank failed'}): "
              "falling back to the NCCL exchange", file=sys.stderr, flush=True)
    _probe_result[key] = result
    return result


def int64_array(values: Sequence[int]) -> "ctypes.Array":
    arr = (ctypes.c_int64 * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = int(v)
    return arr
