"""bench.py contract checks that need no GPU: the reference arm's JSON line, and that the product arm refuses to run
without CUDA instead of falling back to the CPU."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _run(*args):
    return subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=600, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    r = _run("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                      # ONE JSON line
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "lightgcn_train_edges_per_sec" and j["unit"] == "edges/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["warmup"] == 1
    assert j["value"] > 0 and j["ms_per_step"] > 0 and j["vs_baseline"] is None and j["data"] == "synthetic"
    assert j["gpu_launches"] == 0 and "workload" in j["config"] and "model" not in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_product_arm_needs_cuda():
    import torch

    if torch.cuda.is_available():
        return
    r = _run("--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert r.returncode != 0 and "no CPU path" in (r.stderr + r.stdout)


def test_multi_gpu_flags_are_accepted_and_documented():
    """The N > 1 knobs the scaling records name (loss layout, exchange, user-row numbering, overlap) parse, carry their
    measured defaults, and an N > 1 run without torchrun is refused with the launch line instead of silently running N = 1."""
    r = _run("--help")
    assert r.returncode == 0
    for flag in ("--loss-layout", "--exchange", "--user-rows", "--overlap-exchange", "--feature-shards", "--gpus", "--steps", "--warmup",
                 "--impl", "--workload"):
        assert flag in r.stdout, flag
    import torch

    if not torch.cuda.is_available():
        r = _run("--gpus", "2", "--workload", "tiny", "--steps", "1", "--warmup", "1", "--exchange", "peer", "--user-rows", "owner",
                 "--overlap-exchange", "auto", "--loss-layout", "rows")
        assert r.returncode != 0 and ("torch.distributed.run" in (r.stderr + r.stdout) or "no CPU path" in (r.stderr + r.stdout))
