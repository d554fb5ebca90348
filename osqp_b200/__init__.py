"""osqp_b200 -- B200-native (sm_100a) linear-algebra backend for the OSQP QP solver.

The product is `algebra/b200` (plain C) + `osqp_b200/csrc` (hand-written CUDA kernels behind
the C-ABI of include/osqp_b200.h), linked under the unchanged OSQP core.  This Python package
is only the host-side mirror of the solver object used by tests and bench.py.
"""
from .interface import OSQP as _OSQPBase, OSQPError, LoadedLibrary  # noqa: F401
from ._lib import load_library, load_kernels, B200LibraryMissing  # noqa: F401
from . import _capi as constants  # noqa: F401


class OSQP(_OSQPBase):
    """OSQP solver object bound to the B200 backend (`precision` = "f64" | "f32")."""

    def __init__(self, precision="f64"):
        super().__init__(load_library(precision))

    def cg_stats(self):
        """(total CG iterations, number of linear solves) since setup; synchronises."""
        import ctypes as C
        it, ns = C.c_longlong(0), C.c_longlong(0)
        self._lib.osqp_b200_cg_stats(C.cast(self._solver, C.c_void_p), C.byref(it), C.byref(ns))
        return it.value, ns.value
