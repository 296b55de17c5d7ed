"""ctypes binding of libgcf.so -- the C-ABI declared in include/gcf.h.

The library is mandatory: there is no eager / CPU fallback.  If libgcf.so has not been built
(`python -m recommendation_b200.build`, or `__graft_entry__.build()`), importing a kernel entry
point raises immediately instead of silently computing somewhere else.
"""
from __future__ import annotations

import ctypes
from ctypes import c_char_p, c_float, c_int32, c_int64, c_size_t, c_uint64, c_void_p, POINTER
from pathlib import Path
from typing import Optional

import os

# GCF_LIB_PATH: load another build of the same sources (tuning sweeps, tools/sweep_infonce_poly.sh); default = the in-tree library
LIB_PATH = Path(os.environ["GCF_LIB_PATH"]) if os.environ.get("GCF_LIB_PATH") else Path(__file__).with_name("libgcf.so")

MAX_ADDENDS = 8
EPILOGUE_NONE, EPILOGUE_L2NORM = 0, 1
NORM_NONE, NORM_SYM, NORM_ROW = 0, 1, 2
BPR_LOG_EPS_SIGMOID, BPR_SOFTPLUS, BPR_RAW_SCORE = 0, 1, 2
REDUCE_MEAN, REDUCE_SUM = 0, 1


class GcfError(RuntimeError):
    """A libgcf entry point returned a non-zero status."""


class CsrStruct(ctypes.Structure):
    """Mirror of gcf_csr_t (include/gcf.h)."""

    _fields_ = [
        ("n_rows", c_int64), ("n_cols", c_int64), ("nnz", c_int64),
        ("row_ptr", c_void_p), ("col_idx", c_void_p), ("vals", c_void_p),
        ("chunk", c_int32), ("n_long", c_int32), ("n_chunks", c_int32),
        ("long_rows", c_void_p), ("long_chunk_ptr", c_void_p), ("chunk_long", c_void_p),
        ("tiles", c_void_p), ("n_tiles", c_int32), ("n_empty", c_int32), ("empty_rows", c_void_p),
        ("nz_row_ptr", c_void_p), ("nz_rows", c_void_p), ("hub_col_idx", c_void_p),
    ]


# name -> (restype, argtypes); must list every symbol of include/gcf.h (tests/test_abi.py checks it)
_SIGNATURES = {
    "gcf_version": (c_char_p, []),
    "gcf_last_error": (c_char_p, []),
    "gcf_degree_count": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "gcf_bipartite_edge_index": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p]),
    "gcf_coo_to_csr_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "gcf_coo_to_csr_stable": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_norm_values": (c_int32, [c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "gcf_scale_csr_values": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gcf_csr_transpose_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int64]),
    "gcf_csr_transpose": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_spmm_workspace_bytes": (c_size_t, [POINTER(CsrStruct), c_int32]),
    "gcf_spmm_counter_offset": (c_size_t, [POINTER(CsrStruct), c_int32]),
    "gcf_spmm_csr_f32": (c_int32, [POINTER(CsrStruct), c_int32, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_int32, c_float, c_float, c_int32, POINTER(c_void_p), POINTER(c_float),
                                   c_void_p, c_size_t, c_int32, c_void_p]),
    "gcf_propagate_fwd": (c_int32, [POINTER(CsrStruct), c_int32, c_int32, c_void_p, POINTER(c_void_p), c_void_p, c_float,
                                    c_void_p, c_size_t, c_void_p]),
    "gcf_propagate_bwd": (c_int32, [POINTER(CsrStruct), c_int32, c_int32, c_void_p, POINTER(c_void_p), c_float,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_propagate_bwd_adam": (c_int32, [POINTER(CsrStruct), c_int32, c_int32, c_void_p, POINTER(c_void_p), c_float,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float,
                                         c_float, c_float, c_int32, c_int64, c_void_p, c_size_t, c_void_p]),
    "gcf_gather_rows": (c_int32, [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "gcf_scatter_add_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "gcf_scatter_add_rows": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32,
                                       c_void_p, c_size_t, c_void_p]),
    "gcf_slices_to_rows": (c_int32, [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p]),
    "gcf_rows_to_slices": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
    "gcf_peer_alloc": (c_int32, [c_size_t, POINTER(c_void_p)]),
    "gcf_peer_free": (c_int32, [c_void_p]),
    "gcf_peer_export": (c_int32, [c_void_p, c_void_p]),
    "gcf_peer_open": (c_int32, [c_void_p, POINTER(c_void_p)]),
    "gcf_peer_close": (c_int32, [c_void_p]),
    "gcf_peer_barrier": (c_int32, [POINTER(c_void_p), c_int32, c_int32, c_uint64, c_void_p]),
    "gcf_peer_gather_cols": (c_int32, [POINTER(c_void_p), c_int32, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "gcf_peer_sum_cols": (c_int32, [POINTER(c_void_p), c_int32, c_int64, c_void_p, c_int64, c_int64, c_int32, c_void_p]),
    "gcf_peer_copy2d": (c_int32, [POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int32, c_int64, c_int64, c_int32, c_int32,
                                  c_void_p]),
    "gcf_peer_copy_blocks": (c_int32, [POINTER(c_void_p), POINTER(c_int64), POINTER(c_int64), c_int32, c_int64, c_void_p, c_int64,
                                       c_int32, c_void_p]),
    "gcf_sample_negatives": (c_int32, [c_uint64, c_uint64, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p,
                                       c_int32, c_void_p, c_void_p]),
    "gcf_sample_negatives_at": (c_int32, [c_uint64, c_uint64, c_int64, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p,
                                          c_int32, c_void_p, c_void_p]),
    "gcf_sample_negatives_pos": (c_int32, [c_uint64, c_uint64, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p]),
    "gcf_philox_keys": (c_int32, [c_int64, c_uint64, c_uint64, c_void_p, c_void_p]),
    "gcf_csr_dropout_values": (c_int32, [c_void_p, c_int64, c_void_p, c_float, c_uint64, c_uint64, c_void_p, c_void_p]),
    "gcf_bpr_workspace_bytes": (c_size_t, [c_int64]),
    "gcf_bpr_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                              c_int32, c_float, c_int32, c_float, c_float, c_float, c_void_p, c_void_p,
                              c_void_p, c_size_t, c_void_p]),
    "gcf_bpr_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                              c_void_p, c_void_p, c_float, c_float, c_float, c_void_p, c_int64, c_void_p, c_int64, c_void_p]),
    "gcf_bpr_coef_from_scores": (c_int32, [c_void_p, c_int64, c_int32, c_float, c_int32, c_void_p, c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
    "gcf_bpr_fwd_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                  c_int32, c_float, c_int32, c_float, c_float, c_float, c_float, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "gcf_scale_by_device_scalar": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p]),
    "gcf_adam_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float,
                                c_float, c_int32, c_int64, c_void_p]),
    "gcf_sgd_momentum_step": (c_int32, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_float, c_float, c_float, c_int32,
                                        c_int32, c_void_p]),
    "gcf_adam_rows_step": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                     c_float, c_float, c_float, c_float, c_float, c_int32, c_int64, c_void_p]),
    "gcf_kmeans_workspace_bytes": (c_size_t, [c_int64, c_int32, c_int32]),
    "gcf_kmeans_lloyd": (c_int32, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_size_t, c_void_p]),
    "gcf_infonce_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int32]),
    "gcf_infonce_fwd": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32, c_float,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_infonce_bwd": (c_int32, [c_void_p, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int32, c_int32, c_float,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "gcf_csr_sample": (c_int32, [POINTER(CsrStruct), POINTER(CsrStruct), c_void_p, c_void_p]),
    "gcf_spgemm_masked": (c_int32, [POINTER(CsrStruct), POINTER(CsrStruct), POINTER(CsrStruct), c_void_p, c_void_p]),
    "gcf_spgemm_workspace_bytes": (c_size_t, [c_int64]),
    "gcf_spgemm_count": (c_int32, [POINTER(CsrStruct), POINTER(CsrStruct), c_int64, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_spgemm_expand": (c_int32, [POINTER(CsrStruct), POINTER(CsrStruct), c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "gcf_text_workspace_bytes": (c_size_t, [c_int64]),
    "gcf_text_count_records": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_text_parse_pairs": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_sort_unique_workspace_bytes": (c_size_t, [c_int64]),
    "gcf_sort_unique_u64": (c_int32, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "gcf_lookup_sorted_u64": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "gcf_text_parse_pairs_words": (c_int32, [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    "gcf_sort_unique_words_workspace_bytes": (c_size_t, [c_int64]),
    "gcf_sort_unique_words": (c_int32, [c_void_p, c_int32, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "gcf_lookup_sorted_words": (c_int32, [c_void_p, c_int32, c_int64, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "gcf_masked_topn": (c_int32, [c_void_p, c_int64, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_float, c_int32, c_void_p,
                                  c_void_p, c_void_p]),
    "gcf_ranking_hits": (c_int32, [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                   c_void_p, c_void_p]),
    "gcf_directau_workspace_bytes": (c_size_t, [c_int64, c_int32]),
    "gcf_directau_fwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_float, c_void_p,
                                   c_void_p, c_size_t, c_void_p]),
    "gcf_directau_bwd": (c_int32, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int32, c_float, c_void_p, c_void_p,
                                   c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
}

_lib: Optional[ctypes.CDLL] = None


def load() -> ctypes.CDLL:
    """Load libgcf.so (once).  Raises RuntimeError when it is missing -- there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -m recommendation_b200.build` (needs nvcc); this package has no CPU / eager fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the .so is stale
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def exported_symbols():
    return sorted(_SIGNATURES)


def last_error() -> str:
    return load().gcf_last_error().decode("utf-8", "replace")


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise GcfError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t) -> Optional[int]:
    """Device (or host) address of a tensor, None for None.

    The kernels are launched on the CURRENT device's current stream (`current_stream()`), so a tensor that lives on
    another GPU would be dereferenced by the wrong device: refuse it here, where every pointer argument passes."""
    if t is None:
        return None
    if t.is_cuda:
        import torch

        cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: select the device "
                               "first (torch.cuda.set_device / `with torch.cuda.device(...)`), libgcf launches on the current one")
    return t.data_ptr()


def current_stream() -> int:
    """cudaStream_t of the current device's current stream (see `ptr` for the device check on the arguments)."""
    import torch

    return torch.cuda.current_stream().cuda_stream


def ptr_array(tensors):
    """Host array of device pointers (NULL for None entries)."""
    arr = (c_void_p * max(len(tensors), 1))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def ptr_values(addresses):
    """Host array of raw device addresses (ints)."""
    arr = (c_void_p * max(len(addresses), 1))()
    for i, a in enumerate(addresses):
        arr[i] = int(a)
    return arr


def float_array(values):
    arr = (c_float * max(len(values), 1))()
    for i, v in enumerate(values):
        arr[i] = float(v)
    return arr
