"""k-means (gcf_kmeans_lloyd) at the NCL cfg2 shapes: 52,643 users / 91,599 items, d = 64, k = 1000, 25 iterations.
Development / profiling tool (CUDA-event timing; also the ncu target for km_assign_kernel)."""
import json, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from recommendation_b200 import kmeans as km

dev = torch.device("cuda", 0)
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
res = {}
for name, n in (("users", 52643), ("items", 91599)):
    torch.manual_seed(0)
    centers = torch.randn(1000, 64, device=dev)
    x = centers[torch.randint(0, 1000, (n,), device=dev)] + 0.5 * torch.randn(n, 64, device=dev)
    km.kmeans(x, 1000, niter=25)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); c, idx, obj = km.kmeans(x, 1000, niter=25); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    flops = 26 * 2.0 * n * 1000 * 64 * 3          # 26 assignment passes, three bf16 K-slabs each
    res[name] = {"n": n, "k": 1000, "d": 64, "niter": 25, "ms": float(np.median(ts)), "tensor_tflops_as_issued": flops / np.median(ts) / 1e9,
                 "objective": obj, "clusters_used": int(torch.unique(idx).numel())}
print(json.dumps(res, indent=1))
