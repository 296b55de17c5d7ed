"""NVLink bandwidth of csrc/peer.cu's 2-D block mover between the GPUs of one box: pull (remote loads) vs push (remote
stores), contiguous rows vs column pieces of a wider row, as a function of the CTAs per block.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29544 tools/peer_bw.py
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch
import torch.distributed as dist


def main():
    from recommendation_b200 import _lib, peer

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    lib, st = _lib.load(), _lib.current_stream()
    rows, d = 5_000_000, 64
    w = d // world                                   # the column piece one rank owns
    slice_pb = peer.PeerBuffer(rows, w, dev)         # [rows, w]  contiguous rows of w floats   (a rank's column slice)
    full_pb = peer.PeerBuffer(rows, d, dev)          # [rows, d]  full-width rows               (what the loss reads / writes)
    slice_pb.tensor.normal_(); full_pb.tensor.normal_()
    local_slice = torch.randn(rows, w, device=dev)
    local_full = torch.randn(rows, d, device=dev)
    others = [g for g in range(world) if g != rank]
    nb = len(others)
    rows64 = peer.int64_array([rows] * nb)
    bytes_remote = nb * rows * w * 4

    def arr(vals):
        a = (ctypes.c_void_p * len(vals))()
        for i, v in enumerate(vals):
            a[i] = v
        return a

    cases = {
        # name: (src pointers, dst pointers, ld_src, ld_dst)
        "pull contiguous slices -> column pieces of local rows": (arr([slice_pb.base[g] for g in others]),
                                                                 arr([local_full.data_ptr() + 4 * g * w for g in others]), w, d),
        "push my slice -> column piece of every peer's rows   ": (arr([local_slice.data_ptr()] * nb),
                                                                 arr([full_pb.base[g] + 4 * rank * w for g in others]), w, d),
        "pull column pieces of peers' rows -> local slices    ": (arr([full_pb.base[g] + 4 * rank * w for g in others]),
                                                                 arr([local_slice.data_ptr()] * nb), d, w),
        "push column pieces of my rows -> peers' slices       ": (arr([local_full.data_ptr() + 4 * g * w for g in others]),
                                                                 arr([slice_pb.base[g] for g in others]), d, w),
    }
    for name, (src, dst, lds, ldd) in cases.items():
        line = []
        for ctas in (0, 8, 16, 32, 64, 128):
            def run():
                _lib.check(lib.gcf_peer_copy2d(src, dst, rows64, nb, lds, ldd, w, ctas, st), "gcf_peer_copy2d")
            for _ in range(2):
                run()
            dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                run()
            b.record(); b.synchronize()
            ms = torch.tensor([a.elapsed_time(b) / 5], device=dev)
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            line.append(f"{ctas or 'auto':>4}: {float(ms):6.3f} ms {bytes_remote / float(ms) / 1e6:6.0f} GB/s")
        if rank == 0:
            print(f"G={world} w={w:2d} floats  {name}  " + " | ".join(line), flush=True)
    # NCCL reference: all-gather of the same slices
    out = torch.empty(world * rows * w, device=dev)
    for _ in range(2):
        dist.all_gather_into_tensor(out, local_slice.view(-1))
    dist.barrier(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        dist.all_gather_into_tensor(out, local_slice.view(-1))
    b.record(); b.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / 5], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"G={world} w={w:2d} floats  NCCL all_gather_into_tensor of the same slices: {float(ms):6.3f} ms {bytes_remote / float(ms) / 1e6:6.0f} GB/s inbound per rank")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
