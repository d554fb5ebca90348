"""Locate and load the in-tree shared libraries of the B200 backend.

There is deliberately no fallback: if the CUDA extension is missing the import of the
product path fails loudly (the CPU oracle lives under oracle/ and is test infrastructure).
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

from .interface import LoadedLibrary

_HERE = Path(__file__).resolve().parent
# B200_LIBDIR: development only -- a directory with alternative builds of BOTH libraries (make LIBDIR=...)
LIBDIR = Path(os.environ["B200_LIBDIR"]).resolve() if os.environ.get("B200_LIBDIR") else _HERE / "lib"
REPO_ROOT = _HERE.parent
HEADER = REPO_ROOT / "include" / "osqp_b200.h"

_cache = {}


class B200LibraryMissing(ImportError):
    pass


def lib_paths(precision="f64"):
    return (LIBDIR / f"libb200_kernels_{precision}.so", LIBDIR / f"libosqp_b200_{precision}.so")


def load_kernels(precision="f64"):
    """ctypes handle of the kernel library (the C-ABI of include/osqp_b200.h)."""
    key = ("kernels", precision)
    if key not in _cache:
        path, _ = lib_paths(precision)
        if os.environ.get("B200_KERNELS_LIB"):       # development: kernel-library variants
            path = Path(os.environ["B200_KERNELS_LIB"])
        if not path.exists():
            raise B200LibraryMissing(
                f"{path} not found: build it with `make {precision}` (or __graft_entry__.build()); "
                "the B200 backend has no CPU fallback")
        _cache[key] = C.CDLL(str(path), mode=C.RTLD_LOCAL)
    return _cache[key]


def load_library(precision="f64"):
    """libosqp (unchanged OSQP core + algebra/b200) as a LoadedLibrary."""
    key = ("osqp", precision)
    if key not in _cache:
        load_kernels(precision)
        _, path = lib_paths(precision)
        if not path.exists():
            raise B200LibraryMissing(
                f"{path} not found: build it with `make {precision}` (or __graft_entry__.build())")
        dtype = np.float64 if precision == "f64" else np.float32
        L = LoadedLibrary(path, dtype)
        L.lib.osqp_b200_cg_stats.argtypes = [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]
        L.lib.osqp_b200_cg_stats.restype = C.c_int
        L.lib.osqp_b200_sizeof.argtypes = [C.c_int]
        L.lib.osqp_b200_sizeof.restype = C.c_int
        _cache[key] = L
    return _cache[key]


def declared_symbols():
    """Names of every function declared in include/osqp_b200.h."""
    import re
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))
