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
