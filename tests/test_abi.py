"""The C-ABI library loads and exports every symbol include/gcf.h declares (no compute calls)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "gcf.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gcf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ["gcf_degree_count", "gcf_norm_values", "gcf_coo_to_csr_stable", "gcf_csr_transpose", "gcf_spmm_csr_f32",
                 "gcf_propagate_fwd", "gcf_propagate_bwd", "gcf_gather_rows", "gcf_scatter_add_rows", "gcf_bpr_fwd",
                 "gcf_bpr_bwd", "gcf_infonce_fwd", "gcf_infonce_bwd", "gcf_directau_fwd", "gcf_directau_bwd",
                 "gcf_sample_negatives", "gcf_version", "gcf_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from recommendation_b200 import _lib

    assert _lib.LIB_PATH.exists(), "libgcf.so not built: run python -m recommendation_b200.build"
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for name in declared_symbols():
        assert hasattr(lib, name), f"{name} declared in include/gcf.h but not exported by libgcf.so"


def test_binding_covers_every_declared_symbol():
    from recommendation_b200 import _lib

    assert sorted(_lib.exported_symbols()) == declared_symbols()
    lib = _lib.load()
    assert lib.gcf_version().decode().startswith("gcf-b200")
    assert lib.gcf_last_error() is not None


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from recommendation_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libgcf.so")
    with pytest.raises(RuntimeError, match="no CPU / eager fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    import torch
    from recommendation_b200 import functional as F_

    with pytest.raises(RuntimeError, match="CUDA"):
        F_.gather_rows(torch.zeros(4, 8), torch.tensor([0, 1]))
