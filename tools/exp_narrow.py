"""cfg5 experiments on one GPU (development tool, not the contract bench):

  1. L2 fetch granularity (cudaLimitMaxL2FetchGranularity) x narrow-row SpMM variants: the feature-sharded / 2-D
     multi-GPU layouts gather 32 B / 64 B rows, where ncu showed ~120 B of DRAM traffic per missing gather.
  2. One rank of the feature-sharded trainers (world-size-1 process group, d/G-wide tables) under the same knob.
  3. L2 access-policy window (persisting) over the hub rows of X for the d = 64 SpMM, user-row / item-row halves
     launched separately (the synthetic generator's Zipf rank IS the node id, so hubs are contiguous).
"""
import ctypes
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import functional as F_, synth  # noqa: E402
from recommendation_b200.graph import CSRGraph  # noqa: E402

cudart = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so.12")  # same primary context as torch's own runtime copy
LIMIT_FETCH, LIMIT_PERSIST = 0x05, 0x06
ATTR_MAX_PERSIST, ATTR_MAX_WINDOW = 108, 109


def get_limit(which):
    v = ctypes.c_size_t(0)
    rc = cudart.cudaDeviceGetLimit(ctypes.byref(v), which)
    return int(v.value) if rc == 0 else -rc


def set_limit(which, value):
    return cudart.cudaDeviceSetLimit(which, ctypes.c_size_t(value))


def dev_attr(a):
    v = ctypes.c_int(0)
    cudart.cudaDeviceGetAttribute(ctypes.byref(v), a, 0)
    return int(v.value)


class AccessPolicyWindow(ctypes.Structure):
    _fields_ = [("base_ptr", ctypes.c_void_p), ("num_bytes", ctypes.c_size_t), ("hitRatio", ctypes.c_float),
                ("hitProp", ctypes.c_int), ("missProp", ctypes.c_int)]


class StreamAttrValue(ctypes.Union):
    _fields_ = [("pad", ctypes.c_char * 64), ("window", AccessPolicyWindow)]


PROP_NORMAL, PROP_STREAMING, PROP_PERSISTING = 0, 1, 2


def set_window(stream, base, nbytes, ratio, hit=PROP_PERSISTING, miss=PROP_STREAMING):
    v = StreamAttrValue()
    v.window.base_ptr = base
    v.window.num_bytes = nbytes
    v.window.hitRatio = ratio
    v.window.hitProp = hit
    v.window.missProp = miss
    return cudart.cudaStreamSetAttribute(ctypes.c_void_p(stream), 1, ctypes.byref(v))


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    parts = set(sys.argv[2].split(",")) if len(sys.argv) > 2 else {"gran", "rank", "window"}
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    U, I, E, d, K = synth.CONFIGS[cfg]
    if E > 20_000_000:
        users, items = synth.power_law_bipartite_torch(U, I, E, seed=1005, device=dev)
    else:
        inter = synth.power_law_bipartite(U, I, E, seed=1001)
        users, items = torch.from_numpy(inter.users).to(dev), torch.from_numpy(inter.items).to(dev)
    n = U + I
    res = {"cfg": cfg, "limits": {"fetch_default": get_limit(LIMIT_FETCH), "persist_default": get_limit(LIMIT_PERSIST),
                                  "max_persist": dev_attr(ATTR_MAX_PERSIST), "max_window": dev_attr(ATTR_MAX_WINDOW)}}
    print(res["limits"], flush=True)
    out = Path("gpurun_out"); out.mkdir(exist_ok=True)

    def dump():
        (out / f"exp_narrow_{cfg}.json").write_text(json.dumps(res, indent=1))

    if "gran" in parts:
        g = CSRGraph.from_pairs(users, items, U, I, norm="sym")
        res["gran"] = {}
        for gran in (0, 32, 64, 128):
            if gran:
                rc = set_limit(LIMIT_FETCH, gran)
            eff = get_limit(LIMIT_FETCH)
            for dd in (8, 16, 32, 64):
                x = torch.randn(n, dd, device=dev); y = torch.empty_like(x)
                for variant in ((0, 1, 2, 3) if dd <= 16 else (0, 1, 2) if dd == 32 else (0,)):
                    ms = timeit(lambda: F_.spmm_raw(g, x, y=y, variant=variant))
                    if gran == 0:   # every variant must reproduce variant 0 bit for bit (same FMA order)
                        if variant == 0:
                            y_ref = y.clone()
                        else:
                            same = bool(torch.equal(y, y_ref))
                            res["gran"][f"d{dd}_v{variant}_bit_identical"] = same
                            print(f"d={dd} variant={variant} bit-identical to variant 0: {same}", flush=True)
                    res["gran"][f"gran{gran}_eff{eff}_d{dd}_v{variant}"] = ms
                    print(f"gran={gran} (eff {eff}) d={dd} variant={variant}: {ms:.3f} ms", flush=True)
                del x, y
            dump()
        del g
        torch.cuda.empty_cache()

    if "rank" in parts:
        import torch.distributed as dist
        from recommendation_b200.dist import FeatureShardedLightGCNTrainer
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29581")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        res["rank"] = {}
        for dd in (8, 16):
            tr = FeatureShardedLightGCNTrainer(users, items, U, I, d=dd, n_layers=K, seed=1)
            for gran in (128, 64, 32):
                set_limit(LIMIT_FETCH, gran)
                for _ in range(2):
                    tr.step()
                torch.cuda.synchronize()
                marks = []
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(4):
                    tr.step(marks=marks)
                e.record(); torch.cuda.synchronize()
                step = s.elapsed_time(e) / 4
                fwd = np.mean([a.elapsed_time(b) for a, b, _ in marks[0::2]])
                bwd = np.mean([a.elapsed_time(b) for a, b, _ in marks[1::2]])
                res["rank"][f"d{dd}_gran{gran}"] = {"step_ms": step, "fwd_prop_ms": float(fwd), "bwd_prop_adam_ms": float(bwd),
                                                    "loss_part_ms": step - float(fwd) - float(bwd)}
                print(f"one rank of features layout, d/G={dd}, gran={gran}: step {step:.2f} ms (fwd prop {fwd:.2f}, "
                      f"bwd prop+adam {bwd:.2f}, sampler+loss {step-fwd-bwd:.2f})", flush=True)
            del tr
            torch.cuda.empty_cache()
            dump()
        dist.destroy_process_group()

    if "window" in parts:
        set_limit(LIMIT_FETCH, res["limits"]["fetch_default"] or 64)
        res["window"] = {}
        u64, i64 = users.to(torch.int64), items.to(torch.int64)
        a_ui = CSRGraph.from_coo(u64, i64 + U, None, U, n, norm="none")   # user rows gather item rows of X
        a_iu = CSRGraph.from_coo(i64, u64, None, I, n, norm="none")       # item rows gather user rows of X
        x = torch.randn(n, d, device=dev); y = torch.empty_like(x)
        stream = torch.cuda.Stream()
        max_persist = res["limits"]["max_persist"]
        row_bytes = d * 4
        with torch.cuda.stream(stream):
            base_ui = timeit(lambda: F_.spmm_raw(a_ui, x, y=y[:U]))
            base_iu = timeit(lambda: F_.spmm_raw(a_iu, x, y=y[U:]))
            res["window"]["base_ui"], res["window"]["base_iu"] = base_ui, base_iu
            print(f"no window: user rows {base_ui:.3f} ms, item rows {base_iu:.3f} ms", flush=True)
            for persist_mb in (32, 64, 96):
                pb = min(persist_mb << 20, max_persist)
                rc0 = set_limit(LIMIT_PERSIST, pb)
                for win_mb, ratio in ((persist_mb, 1.0), (2 * persist_mb, 0.5), (persist_mb // 2, 1.0)):
                    wb = min(win_mb << 20, res["limits"]["max_window"])
                    for miss in (PROP_STREAMING, PROP_NORMAL):
                        rc1 = set_window(stream.cuda_stream, x.data_ptr() + U * row_bytes, wb, ratio, PROP_PERSISTING, miss)
                        t_ui = timeit(lambda: F_.spmm_raw(a_ui, x, y=y[:U]))
                        rc2 = set_window(stream.cuda_stream, x.data_ptr(), wb, ratio, PROP_PERSISTING, miss)
                        t_iu = timeit(lambda: F_.spmm_raw(a_iu, x, y=y[U:]))
                        key = f"persist{persist_mb}_win{win_mb}_r{ratio}_miss{miss}"
                        res["window"][key] = {"ui": t_ui, "iu": t_iu, "rc": [rc0, rc1, rc2]}
                        print(f"{key}: user rows {t_ui:.3f} ms, item rows {t_iu:.3f} ms (rc {rc0},{rc1},{rc2})", flush=True)
                dump()
            set_window(stream.cuda_stream, 0, 0, 0.0, PROP_NORMAL, PROP_NORMAL)
            cudart.cudaCtxResetPersistingL2Cache()
    dump()


if __name__ == "__main__":
    main()
