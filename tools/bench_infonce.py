"""Micro-benchmark of the tcgen05 InfoNCE kernels at the cfg2 (NCL / SSL4Rec) and cfg3 (DirectAU) shapes.
Development / profiling tool; prints one JSON line per case (CUDA-event timing, warm)."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import _lib, functional as F_

PEAK_TF = 1395.7  # bf16 sustained, MEASURED_PEAKS.json


def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda", 0)
    lib = _lib.load()
    cases = [("ssl_layer users  B x U", 4096, 52643, 64), ("ssl_layer items  B x I", 4096, 91599, 64), ("InfoNCE in-batch  B x B", 4096, 4096, 64),
             ("SSL4Rec tower out B x B", 4096, 4096, 128), ("gcl users U x U", 52643, 52643, 64), ("DirectAU Gram B x B", 2048, 2048, 128)]
    only = sys.argv[1] if len(sys.argv) > 1 else None
    for name, m, n, d in cases:
        if only and only not in name:
            continue
        q = torch.randn(m, d, device=dev); k = torch.randn(n, d, device=dev)
        pos_idx = torch.randint(0, n, (m,), device=dev)
        ws_bytes = lib.gcf_infonce_workspace_bytes(m, n, d); ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        row = torch.empty(m, device=dev); pos = torch.empty(m, device=dev)
        st = _lib.current_stream()
        fwd = lambda: _lib.check(lib.gcf_infonce_fwd(_lib.ptr(q), d, m, _lib.ptr(k), d, n, d, 1, 0.2, _lib.ptr(pos_idx), _lib.ptr(row), None,
                                                     _lib.ptr(pos), _lib.ptr(ws), ws_bytes, st), "fwd")
        t_f = timeit(fwd)
        w_row = torch.full((m,), 1.0 / m, device=dev); w_pos = -w_row
        gq, gk = torch.empty_like(q), torch.empty_like(k)
        bwd = lambda: _lib.check(lib.gcf_infonce_bwd(_lib.ptr(q), d, m, _lib.ptr(k), d, n, d, 1, 0.2, _lib.ptr(pos_idx), _lib.ptr(row), None,
                                                     _lib.ptr(w_row), None, _lib.ptr(w_pos), _lib.ptr(gq), d, _lib.ptr(gk), d,
                                                     _lib.ptr(ws), ws_bytes, st), "bwd")
        t_b = timeit(bwd)
        f_f, f_b = 2.0 * m * n * d, 8.0 * m * n * d
        # fp32 eager reference timing of the same math on the same GPU (what the reference's op chain costs here)
        def eager():
            s = torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(k, dim=1).T / 0.2
            return torch.logsumexp(s, 1)
        t_e = timeit(eager, iters=3, warm=1) if m * n <= 4096 * 100000 else None
        print(json.dumps({"case": name, "M": m, "N": n, "d": d, "fwd_ms": t_f, "fwd_tflops": f_f / t_f / 1e9, "fwd_frac_of_bf16_peak": f_f / t_f / 1e9 / PEAK_TF,
                          "bwd_ms": t_b, "bwd_tflops_as_implemented": f_b / t_b / 1e9, "bwd_frac_of_bf16_peak": f_b / t_b / 1e9 / PEAK_TF,
                          "logits_per_ns_fwd": m * n / t_f / 1e6, "eager_fp32_lse_ms": t_e}), flush=True)


if __name__ == "__main__":
    main()
