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


def _prototypes():
    """{name: (return type, [parameter types])} parsed from include/gcf.h."""
    text = (ROOT / "include" / "gcf.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"^\s*#.*$", "", text, flags=re.M)
    out = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(gcf_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [" ".join(p.split()) for p in params.split(",")]
        out[name] = (ret, plist)
    return out


def _ctype_of(decl: str):
    """ctypes type the binding must use for a C parameter declaration (name stripped)."""
    from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p

    if "*" in decl:
        return "pointer"
    base = decl.rsplit(" ", 1)[0] if " " in decl else decl
    base = base.replace("const ", "").strip()
    return {"int32_t": c_int32, "int": c_int32, "int64_t": c_int64, "uint64_t": c_uint64, "size_t": c_size_t, "float": c_float,
            "gcf_stream_t": c_void_p, "const char": c_char_p}.get(base, base)


def test_binding_signatures_match_the_header():
    """Every ctypes signature in _lib.py has the arity of its prototype in include/gcf.h, scalar parameters use the matching
    ctypes type and pointer parameters a pointer type (a mismatch would silently corrupt the call frame)."""
    import ctypes as C
    from recommendation_b200 import _lib

    protos = _prototypes()
    assert sorted(protos) == declared_symbols()
    for name, (restype, argtypes) in _lib._SIGNATURES.items():
        ret, params = protos[name]
        assert len(argtypes) == len(params), f"{name}: binding has {len(argtypes)} arguments, header declares {len(params)}"
        for k, (decl, at) in enumerate(zip(params, argtypes)):
            want = _ctype_of(decl)
            if want == "pointer":
                is_ptr = at in (C.c_void_p, C.c_char_p) or hasattr(at, "contents") or getattr(at, "_type_", None) is not None and issubclass(at, C._Pointer)
                assert is_ptr, f"{name} argument {k} ({decl}) must be bound as a pointer, not {at}"
            else:
                assert at is want, f"{name} argument {k} ({decl}) is bound as {at}, expected {want}"
        if ret in ("int", "int32_t"):
            assert restype is C.c_int32, name
        elif ret == "size_t":
            assert restype is C.c_size_t, name
        elif "char" in ret:
            assert restype is C.c_char_p, name
