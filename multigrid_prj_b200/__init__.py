"""multigrid_prj_b200 -- B200-native multigrid solve phase (GMG + AMG) behind a C ABI.

The product is `lib/libmgb200.so` (hand-written sm_100a CUDA, built in-tree by
`multigrid_prj_b200.build`); its entry points are declared in `include/mgb200.h`.
The C++ facade under `dropin/` re-declares the reference's solver classes on top of that ABI so
the reference's own drivers compile against it.  This Python package is plumbing only: a ctypes
loader and thin handles used by the tests and by bench.py.  There is no CPU fallback.
"""
from ._lib import load, MgbError, lib_path   # noqa: F401
from .gmg import GmgConfig, Gmg              # noqa: F401
from .amg import Amg, System                 # noqa: F401

__all__ = ["load", "MgbError", "lib_path", "GmgConfig", "Gmg", "Amg", "System"]
