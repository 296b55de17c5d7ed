#!/usr/bin/env python
"""bench.py -- LightGCN training throughput (edges/sec) on B200, the BASELINE.json headline metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg1|cfg5|...] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

One "step" = one full-batch LightGCN optimisation step (lightgcn.py:83-120) over all E training interactions of a
synthetic power-law bipartite graph: K-layer propagation -> Philox negatives -> fused gather+BPR+reg ->
backward (BPR scatter, transpose propagation) -> Adam.  `value` = E * steps / time (whole job).

  value     : device-resident inputs, the fixed libgcf launch sequence (FusedLightGCNTrainer / ShardedLightGCNTrainer)
  e2e       : the same step through the reference-facing API (LightGCN.forward + bpr_step_loss + backward +
              optimizer.step()) with the step's index tensors in PINNED HOST memory: H2D copy and the D2H read of
              the loss are inside the timed region
  roofline  : SpMM (dominant kernel) algorithmic bytes / average launch duration, measured with CUDA events
              inside the timed region, against the measured HBM peak of MEASURED_PEAKS.json
  cpu_baseline / --impl reference : the oracle's CPU restatement of lightgcn.py's step (torch, all host cores)
              on a bounded sample -- the only place outside tests/ and smoke() that executes oracle/.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

METRIC = "lightgcn_train_edges_per_sec"
UNIT = "edges/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def measured_peak():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload: str):
    """DRAM bytes per SpMM launch from the committed `ncu --set full` capture of this workload (profiles/traffic.json),
    or None when no capture exists for it."""
    p = ROOT / "profiles" / "traffic.json"
    try:
        rec = json.loads(p.read_text())[workload]["spmm_csr_kernel"]
        return float(rec["dram_bytes_per_launch"])
    except Exception:
        return None


def traffic_source(workload: str) -> str:
    try:
        rec = json.loads((ROOT / "profiles" / "traffic.json").read_text())[workload]["spmm_csr_kernel"]
        return (f"profiles/traffic.json <- {rec.get('source', '?')} (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                f"a committed capture of {rec.get('kernel', 'the SpMM kernel')}, not re-measured by this run)")
    except Exception:
        return "none"


def spmm_algorithmic_bytes(n_rows, n_cols, nnz, d):
    """nnz*(4+4) + (N_r+1)*4 + N_c*d*4 + N_r*d*4 per SpMM launch (compulsory traffic)."""
    return nnz * 8 + (n_rows + 1) * 4 + n_cols * d * 4 + n_rows * d * 4


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 100 ms in a side process; samples are selected by wall-clock window."""

    FIELDS = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="gcf_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self, t_begin: float, t_end: float):
        """Summarise the samples taken in the wall-clock window [t_begin, t_end] (time.time() values)."""
        import datetime

        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        clocks, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in Path(self.path).read_text().splitlines():
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                    clk = float(f[2]); mx_here = float(f[3])
                except ValueError:
                    continue
                if ts < t_begin - 0.05 or ts > t_end + 0.05:
                    continue
                clocks.append(clk); mx = mx_here
                for nm, val in zip(names, f[6:10]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        if clocks:
            out = {"sm_mhz": statistics.median(clocks), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(clocks)}
        return out


# ------------------------------------------------------------------------------------------------- CPU arm
def cpu_sample_shape(workload: str):
    """Bounded CPU sample of a workload: the graph itself when it has <= 2.1 M edges, otherwise the same generator at
    1/k scale (same users:items:edges proportions, hence the same average degrees), k chosen so that E ~ 5 M (tables of
    750 k x 64 floats = 192 MB: well outside the host caches; ~5 s per step on 16 threads, so that the reference arm can run
    the driver's own --steps / --warmup within a few minutes)."""
    from recommendation_b200 import synth

    U, I, E, d, K = synth.CONFIGS[workload]
    if E <= 5_250_000:
        return U, I, E, d, K, f"the {workload} graph"
    k = max(1, round(E / 5_000_000))
    return U // k, I // k, E // k, d, K, f"a 1/{k}-scale {workload} graph (same generator and degree law)"


def cpu_reference_step_rate(workload: str, steps: int, warmup: int):
    """Time the oracle's restatement of lightgcn.py's epoch iteration on the host cores (bounded sample)."""
    import torch
    from oracle import lightgcn_ref
    from recommendation_b200 import synth
    from recommendation_b200.lightgcn import build_edge_index

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    U, I, E, d, K, what = cpu_sample_shape(workload)
    inter = synth.power_law_bipartite(U, I, E, seed=1001)
    pu, pi = torch.from_numpy(inter.users), torch.from_numpy(inter.items)
    ei = build_edge_index(pu, pi, U)
    torch.manual_seed(1)
    uw = torch.nn.init.xavier_uniform_(torch.empty(U, d)).requires_grad_(True)
    iw = torch.nn.init.xavier_uniform_(torch.empty(I, d)).requires_grad_(True)
    opt = torch.optim.Adam([uw, iw], lr=0.01)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad()
        neg = torch.randint(0, I, (E,))
        loss = lightgcn_ref.lightgcn_step_loss(uw, iw, ei, pu, pi, neg, K, 1e-4)
        loss.backward()
        opt.step()
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    per_step = sum(times) / len(times)
    return E / per_step, per_step, cores, (f"{steps} full step(s) of lightgcn.py's epoch loop (PyG LGConv restated, oracle/lightgcn_ref.py) on "
                                           f"{what}: U={U} I={I} E={E} d={d} K={K}, torch CPU fp32, {cores} threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, args.steps), max(0, args.warmup)   # the driver's own K and W, on the bounded sample
    value, per_step, cores, sample = cpu_reference_step_rate(args.workload, steps, warmup)
    U, I, E, d, K = __import__("recommendation_b200.synth", fromlist=["CONFIGS"]).CONFIGS[args.workload]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: LightGCN {K}-layer d={d} full-batch BPR, U={U} I={I} E={E}",
                   "reference_sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- GPU arm
def make_workload(name: str, device, seed=None):
    import torch
    from recommendation_b200 import synth

    U, I, E, d, K = synth.CONFIGS[name]
    if seed is None:
        seed = 1000 + (int(name[3:]) if name.startswith("cfg") else 0)
    if E >= 20_000_000:
        users, items = synth.power_law_bipartite_torch(U, I, E, seed=seed, device=device)
    else:
        inter = synth.power_law_bipartite(U, I, E, seed=seed)
        users, items = torch.from_numpy(inter.users).to(device), torch.from_numpy(inter.items).to(device)
    return U, I, E, d, K, users, items


def measure_secondary(name: str, dev, steps: int, warmup: int, peak: float):
    """Device-resident step time + SpMM roofline of a second, small workload (cfg1: the graph BASELINE.json's >= 60 % target
    is quoted on), reported inside the same JSON line under "secondary" so that the driver's record holds it."""
    import torch
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer
    from recommendation_b200.tables import xavier_uniform_table

    U, I, E, d, K, users, items = make_workload(name, dev)
    n = U + I
    graph = CSRGraph.from_pairs(users, items, U, I, norm="sym")
    trainer = FusedLightGCNTrainer(graph, U, I, xavier_uniform_table(U, I, d, seed=1234, device=dev), users, items, n_layers=K,
                                   lr=0.01, reg_weight=1e-4, seed=1234)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(warmup):
        trainer.step()
    torch.cuda.synchronize()
    step_ms, spmm_us = [], []
    for _ in range(steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = []
        s.record(); loss = trainer.step(marks=marks); e.record(); e.synchronize()
        step_ms.append(s.elapsed_time(e))
        a, b, launches = marks[0]
        spmm_us.append(a.elapsed_time(b) * 1e3 / launches)
    ms = sum(step_ms) / len(step_ms)
    us = sum(spmm_us) / len(spmm_us)
    alg = spmm_algorithmic_bytes(n, n, graph.nnz, d) + n * d * 4 // K
    traffic = ncu_traffic(name)
    return {"workload": f"{name}: LightGCN {K}-layer d={d} full-batch BPR + Adam, U={U} I={I} E={E} (nnz={graph.nnz})",
            "value": E / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
            "l2": "flushed between timed steps (256 MiB write), each step timed with its own CUDA-event pair",
            "loss_final": float(loss.item()),
            "roofline": {"bound": "hbm", "kernel": "spmm_flat_kernel", "achieved": alg / (us * 1e-6) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (us * 1e-6) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes_per_launch": alg,
                         "avg_launch_us": us,
                         "note": "the tables (18 MB) and the CSR (16 MB) are L2-resident within a step: the launch is bound by the "
                                 "L2 -> SM path (526 MB of row gathers per launch), not by HBM; reported against the HBM peak as the contract asks"}}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from recommendation_b200 import _lib
    from recommendation_b200.graph import CSRGraph
    from recommendation_b200.lightgcn import FusedLightGCNTrainer, LightGCN, build_edge_index
    from recommendation_b200.tables import xavier_uniform_table

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N>1 must be launched with torch.distributed.run --nproc-per-node N")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices: recommendation_b200 has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.load()

    U, I, E, d, K, users, items = make_workload(args.workload, dev)
    n = U + I
    torch.manual_seed(1)
    peak, peak_src = measured_peak()
    flush = None
    working_set_small = E < 20_000_000
    if working_set_small:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # 2x the 126 MB L2

    if world == 1:
        graph = CSRGraph.from_pairs(users, items, U, I, norm="sym")
        table = xavier_uniform_table(U, I, d, seed=1234, device=dev)   # the same values at every N (tables.py)
        trainer = FusedLightGCNTrainer(graph, U, I, table, users, items, n_layers=K, lr=0.01, reg_weight=1e-4, seed=1234)
        nnz, spmm_rows, spmm_d = graph.nnz, n, d
        parallelism = "single GPU"
        scaling = "strong"   # the graph (total work) is the same for every N: the N > 1 runs shard THIS workload
    else:
        from recommendation_b200.dist import FeatureShardedLightGCNTrainer, ShardedLightGCNTrainer

        # layout: F feature shards x R = G / F row shards.  F = G: no collective in the propagation at all (one all-reduce
        # of E floats per step); F = 1: the row-sharded layout (one all-gather of d-wide rows per layer); in between: the
        # all-gathers run inside row groups on d/F-wide rows.  Default picked from the measured r01 sweep.
        F = args.feature_shards
        if F <= 0:
            F = {2: 2, 4: 4, 8: 8}.get(world, 1)
        while F > 1 and (world % F != 0 or d % F != 0 or (d // F) % 4 != 0 or d // F < 8):
            F //= 2
        if args.loss_layout == "auto":
            args.loss_layout = "rows"   # r03: with the peer-memory exchange the rows layout wins at every N (2 GPUs 49.7 vs 54.5 ms)
        if F == world:
            trainer = FeatureShardedLightGCNTrainer(users, items, U, I, d=d, n_layers=K, lr=0.01, reg_weight=1e-4, seed=1234,
                                                    loss_layout=args.loss_layout, exchange=args.exchange, user_rows=args.user_rows,
                                                    overlap=(args.overlap_exchange == "on" or
                                                             (args.overlap_exchange == "auto" and args.exchange == "peer")))
            nnz, spmm_rows, spmm_d = trainer.local_nnz, n, trainer.dg
            parallelism = (f"feature-sharded over {world} GPUs: every rank owns d/G = {trainer.dg} columns of all [N, d] tables, "
                           f"propagation / Adam without any collective; ")
            if args.loss_layout == "rows":
                if trainer.exchange == "peer":
                    parallelism += ("loss on full-width rows of one user block per rank: every rank pulls the column slices it needs "
                                    "straight out of the peers' memory over NVLink (one gather kernel per table that also lays the "
                                    "rows out), fused BPR on E/G triples, gradients pulled back the same way (one summing kernel for "
                                    "the items, one copy kernel for the users); device-side barriers over peer flag words, no NCCL "
                                    "collective on the data path"
                                    + ("; the item-side transfers run on a high-priority side stream next to the user-row block of "
                                       "the last forward layer / the item-row block of the first backward product"
                                       if trainer.overlap else ""))
                else:
                    parallelism += ("loss on full-width rows of one user block per rank: item slices all-gathered, user slices "
                                "all-to-all'ed, fused BPR on E/G triples, gradients returned by one reduce-scatter + one all-to-all"
                                + ("; exchanges overlapped with item-row / user-row blocks of the adjacent propagation layers"
                                   if trainer.overlap else ""))
            else:
                parallelism += "ONE NCCL all-reduce of E fp32 partial scores per step"
        else:
            trainer = ShardedLightGCNTrainer(users, items, U, I, d=d, n_layers=K, lr=0.01, reg_weight=1e-4, seed=1234,
                                             feature_shards=F)
            nnz, spmm_rows, spmm_d = trainer.local_nnz, trainer.rows_per_rank, trainer.d
            if F == 1:
                parallelism = f"row-sharded tables + adjacency rows over {world} GPUs, one NCCL all-gather per propagation layer"
            else:
                parallelism = (f"2-D: {world // F} row shards x {F} feature shards; tables + adjacency rows row-sharded inside each row "
                               f"group (one NCCL all-gather of d/{F}-wide rows per propagation layer), BPR scores completed by one "
                               f"all-reduce of E/{world // F} floats inside each feature group")
        scaling = "strong"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    loss_warm = None
    for _ in range(args.warmup):
        loss_warm = trainer.step()
    loss_after_warmup = float(loss_warm.item()) if loss_warm is not None else None
    barrier()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
        time.sleep(0.5)  # let nvidia-smi come up before the timed region
    step_ms, spmm_us = [], []
    ev = lambda: torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.perf_counter(); clk_t0 = time.time()
    for _ in range(args.steps):
        if flush is not None:
            flush.zero_()
        s, e = ev(), ev()
        marks = []
        s.record()
        loss_last = trainer.step(marks=marks)
        e.record()
        e.synchronize()
        step_ms.append(s.elapsed_time(e))
        # (start event, end event, number of SpMM launches in between); the first pair brackets the K forward launches, the
        # second the K backward ones -- whose last launch also carries the fused Adam epilogue on one GPU, so only the
        # forward pair is a pure SpMM timing there
        pure = marks[:1] if getattr(trainer, "fused_adam", False) else marks
        for (a, b, launches) in pure:
            spmm_us.append(a.elapsed_time(b) * 1e3 / launches)
    barrier()
    wall = time.perf_counter() - wall0; clk_t1 = time.time()
    # self-check of the run (outside the timed region): the loss of the last warm-up and of the last timed step and a
    # checksum of the trained table.  Initial table, triples and Philox negatives are functions of the seed alone, so these
    # agree across N = 1/2/4/8 up to summation order.
    loss_final = float(loss_last.item())
    csum = trainer.table.double().abs().sum()
    if world > 1:
        dist.all_reduce(csum, op=dist.ReduceOp.SUM)
    check = {"loss_after_warmup": loss_after_warmup, "loss_final": loss_final, "table_abs_sum": float(csum.item()),
             "steps_run": args.warmup + args.steps,
             "note": "same seed => same initial table, triples and Philox negatives at every N; values agree across N up to fp32 summation order"}
    clk_note = "sampled during the timed region"
    if wall < 1.0:
        # the timed region is shorter than a few sampler periods: keep the identical step loop running (untimed)
        # for ~1 s right after it and sample the clocks over that window instead
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            for _ in range(10):
                if flush is not None:
                    flush.zero_()
                trainer.step()
            torch.cuda.synchronize()
        clk_t1 = time.time()
        clk_note = f"timed region {wall * 1e3:.1f} ms is shorter than the sampler period; sampled over it plus 1 s of the identical step loop run right after"
    clk = clocks.stop(clk_t0, clk_t1) if rank == 0 else None
    if clk is not None:
        clk["note"] = clk_note

    total_ms = sum(step_ms)
    launches_per_step = trainer.launches_per_step
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = E * args.steps / (total_ms * 1e-3)
    spmm_avg_us = sum(spmm_us) / max(len(spmm_us), 1)
    alg_bytes = spmm_algorithmic_bytes(spmm_rows, n, nnz, spmm_d)  # per launch on ONE rank
    alg_note = "SURVEY 8(d) B_spmm = nnz*8 + (N_r+1)*4 + N_c*d*4 + N_r*d*4"
    if getattr(trainer, "fused_adam", False):
        # the timed pair brackets the whole forward propagation, whose last launch also writes the layer sum:
        # SURVEY 8(d) "propagate fwd (K layers + mean) = K * B_spmm + N*d*4", averaged over its K launches
        alg_bytes += spmm_rows * spmm_d * 4 // K
        alg_note = "SURVEY 8(d) forward propagation (K * B_spmm + N_r*d*4) / K launches"
    achieved = alg_bytes / (spmm_avg_us * 1e-6) / 1e9 if spmm_avg_us > 0 else 0.0
    traffic = ncu_traffic(args.workload) if world == 1 else None  # DRAM bytes actually moved per launch (ncu capture)
    kernel_name = "spmm_flat_kernel" if spmm_d >= 32 else "spmm_csr_kernel"   # csrc/spmm.cu: dispatch by row width

    # ---- e2e: the same step through the public API with HOST index buffers; H2D of the step's inputs and the D2H read of
    # the loss are inside the timed region.  N = 1: the reference-facing classes (LightGCN.forward + bpr_step_loss + autograd
    # + torch.optim.Adam); N > 1: the sharded trainer's step() fed from rank-local pinned index shards.
    e2e = None
    if not args.no_e2e and world == 1:
        from recommendation_b200 import functional as F_
        from recommendation_b200.lightgcn import bpr_step_loss

        del trainer  # free the device-resident arm's buffers (cfg5: ~35 GB) before the public-API arm allocates its own
        torch.cuda.empty_cache()
        torch.manual_seed(1)
        model = LightGCN(U, I, d, K).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=0.01, fused=True)
        ei = build_edge_index(users, items, U)
        # the interaction list is user-major in the host buffers (a one-off data-preparation step outside the timed region:
        # lightgcn.py:86-88 consumes all interactions of an epoch as one batch, so their order is free); the fused BPR
        # kernel then keeps each user's row and gradient in registers over the run of its interactions
        order = torch.argsort(users * I + items)
        # int32 on the host and on the wire (node ids < 2^31), widened on the device by the API's own index normalisation
        pu_h, pi_h = users[order].to(torch.int32).cpu().pin_memory(), items[order].to(torch.int32).cpu().pin_memory()
        del order
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()
        model(ei)  # builds + caches the CSR (the reference normalises on every call; here once)
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        e2e_steps = max(1, min(args.steps, 10))
        for it in range(2 + e2e_steps):
            if it == 2:
                torch.cuda.synchronize(); t0 = time.perf_counter()
            # the step's index tensors come from pinned host memory on a copy stream; the propagation (which does not
            # need them) runs on the main stream meanwhile
            with torch.cuda.stream(copy_stream):
                pu_d = pu_h.to(dev, non_blocking=True); pi_d = pi_h.to(dev, non_blocking=True)
                copied = torch.cuda.Event(); copied.record(copy_stream)
            opt.zero_grad(set_to_none=True)
            user_emb, item_emb = model(ei)
            neg = F_.sample_negatives(E, I, seed=1234, offset=it, device=dev)
            main.wait_event(copied)
            pu_d.record_stream(main); pi_d.record_stream(main)
            loss = bpr_step_loss(user_emb, item_emb, pu_d, pi_d, neg, 1e-4)
            loss.backward()
            opt.step()
            loss_h.copy_(loss.detach(), non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / e2e_steps
        e2e = {"value": E / dt, "unit": UNIT, "h2d_bytes_per_step": int(pu_h.numel() * pu_h.element_size() * 2), "d2h_bytes_per_step": 4,
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "loss_last": float(loss_h.item()),
               "api": "LightGCN.forward(edge_index) + sample_negatives + bpr_step_loss + loss.backward() + torch.optim.Adam(fused=True).step(); "
                      "index tensors (user-major interaction list, int32) copied from pinned host memory on a side stream every step, loss read back to the host; "
                      "the normalised CSR is built once by the first forward call, OUTSIDE the timed region (one-off, 171 ms at cfg5; the "
                      "reference re-normalises inside every LGConv call)"}
        del model, opt
    elif not args.no_e2e:
        # rank-local index shards in pinned host memory -> H2D on a side stream -> trainer.step() -> loss D2H, every step
        bufs = trainer.index_buffers()
        host = [b.to(torch.int32).cpu().pin_memory() for b in bufs]       # int32 on the host and on the wire (ids < 2^31)
        stage = [torch.empty(b.shape, dtype=torch.int32, device=dev) for b in bufs]
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        e2e_steps = max(1, min(args.steps, 10))
        t_sum = 0.0
        for it in range(2 + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_stream(main)      # the previous step has finished reading the index buffers
                for b, h, sg in zip(bufs, host, stage):
                    sg.copy_(h, non_blocking=True)
                    b.copy_(sg)                        # widened to the trainer's int64 index buffers on the device
                copied = torch.cuda.Event(); copied.record(copy_stream)
            loss = trainer.step(wait_before_loss=copied)   # the propagation does not need the triples: it overlaps the copy
            loss_h.copy_(loss.detach(), non_blocking=True)
            torch.cuda.synchronize()
            if it >= 2:
                t_sum += time.perf_counter() - t0
        t = torch.tensor([t_sum], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        h2d = torch.tensor([sum(h.numel() * h.element_size() for h in host)], dtype=torch.float64, device=dev)
        dist.all_reduce(h2d, op=dist.ReduceOp.SUM)
        dt = float(t.item()) / e2e_steps
        e2e = {"value": E / dt, "unit": UNIT, "h2d_bytes_per_step": int(h2d.item()), "d2h_bytes_per_step": 4 * world,
               "ms_per_step": dt * 1e3, "steps": e2e_steps, "loss_last": float(loss_h.item()),
               "api": f"{type(trainer).__name__}.step() on every rank; the rank's training-triple index arrays (int32) are copied from rank-local pinned "
                      "host memory on a side stream every step (bytes summed over ranks), the global loss is read back to the host on every "
                      "rank; time = max over ranks of the per-step wall time between a barrier and the loss read; the sharded operator is "
                      "built once outside the timed region"}

    secondary = None
    if world == 1 and args.workload == "cfg5" and not args.no_secondary:
        secondary = measure_secondary("cfg1", dev, 20, 5, peak)

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, per_step, cores, sample = cpu_reference_step_rate(args.workload, 2, 1)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "ms_per_step": per_step * 1e3}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: LightGCN {K}-layer d={d} full-batch BPR + Adam, U={U} I={I} E={E} (nnz={2 * E})",
                       "parallelism": parallelism,
                       "l2": ("flushed between timed steps (256 MiB write), each step timed with its own CUDA-event pair"
                              if flush is not None else "inputs >> 126 MB L2 (tables 3.84 GB, CSR 1.6 GB); steps timed back to back"),
                       "negatives": "Philox4x32-10 on device, uniform without rejection (lightgcn.py:91-94)",
                       "wall_s_timed_region": wall},
            "clocks": clk,
            "check": check,
            "e2e": e2e,
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                         "dram_rate_frac_of_peak": (traffic / (spmm_avg_us * 1e-6) / 1e9 / peak) if traffic and spmm_avg_us > 0 else None,
                         "traffic_source": traffic_source(args.workload),
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "algorithmic_bytes_note": alg_note, "avg_launch_us": spmm_avg_us,
                         "launches_timed": len(spmm_us) * K,
                         "launches_note": "forward-propagation SpMM launches of the timed steps (the last one carries the layer-sum epilogue)"},
            "cpu_baseline": cpu_baseline,
            "secondary": secondary,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", default=os.environ.get("GCF_BENCH_WORKLOAD", "cfg5"))
    ap.add_argument("--loss-layout", choices=("auto", "rows", "scores"), default=os.environ.get("GCF_BENCH_LOSS_LAYOUT", "auto"),
                    help="feature-sharded layout only: where the BPR loss is evaluated (see dist.FeatureShardedLightGCNTrainer); "
                         "auto = measured default: full-width rows at every N")
    ap.add_argument("--exchange", choices=("peer", "nccl"), default=os.environ.get("GCF_BENCH_EXCHANGE", "peer"),
                    help="feature-sharded layout, loss on rows: 'peer' = slices pulled out of the peers' memory over NVLink by "
                         "csrc/peer.cu (default), 'nccl' = all-gather / all-to-all / reduce-scatter + layout passes (r01/r02)")
    ap.add_argument("--user-rows", choices=("owner", "natural"), default=os.environ.get("GCF_BENCH_USER_ROWS", "owner"),
                    help="feature-sharded layout, peer exchange: numbering of the user rows inside the trainer's tables "
                         "(owner = each rank's users contiguous in every slice: contiguous NVLink transfers)")
    ap.add_argument("--overlap-exchange", choices=("auto", "on", "off"), nargs="?", const="on",
                    default=os.environ.get("GCF_BENCH_OVERLAP", "auto"),
                    help="feature-sharded layout, loss on rows: run the item-side exchanges next to row blocks of the adjacent "
                         "propagation layers.  auto = on with the peer-memory exchange (cfg5, 8 GPUs: 25.8 -> 24.2 ms; 4 GPUs: "
                         "36.7 -> 36.0 ms), off with the NCCL exchange (measured neutral in r01)")
    ap.add_argument("--feature-shards", type=int, default=int(os.environ.get("GCF_BENCH_FEATURE_SHARDS", "0")),
                    help="N > 1 only: F feature shards x N/F row shards (1 = row-sharded, N = feature-sharded, 0 = measured default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg1 record that rides in the default (cfg5, N=1) line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3  # timing rule: >= 3 warm-up steps
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
