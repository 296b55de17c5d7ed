"""recommendation_b200 -- B200-native (sm_100a) graph-collaborative-filtering hot path.

A from-scratch implementation of the LightGCN-style propagation, losses and embedding updates shared by
the models of Cmint22/Recommendation, behind that repository's own Python class / function interface.
All arithmetic runs in hand-written CUDA kernels reached through the C-ABI of libgcf.so (include/gcf.h);
importing this package does not need a GPU, calling any kernel does.
"""
from . import _lib
from ._lib import GcfError

__version__ = "0.1.0"


def library_version() -> str:
    return _lib.load().gcf_version().decode()


__all__ = ["GcfError", "library_version", "__version__"]
