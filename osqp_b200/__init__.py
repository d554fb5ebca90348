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

    def solve_batch(self, l, u, q=None):
        """Solve a batch of QPs that share P and A with this (set-up) solver and differ in their bounds
        `l`, `u` (arrays of shape (nb, m)) and optionally their linear cost `q` ((nb, n)): one CTA per QP,
        the whole ADMM loop on the device (osqp_b200/csrc/batch.cu; BASELINE configs[4]).  Returns a
        namespace of arrays: x (nb, n), y (nb, m), iter, status_val, obj_val, prim_res, dual_res,
        cg_iters, rho_updates.  A QP reported as OSQP_MAX_ITER_REACHED (7) has not been checked for
        infeasibility: re-solve it with update(l=..., u=...) + solve()."""
        import ctypes as C
        from types import SimpleNamespace
        import numpy as np
        from . import _capi
        dt = self._L.dtype
        inf = _capi.OSQP_INFTY
        l = np.ascontiguousarray(np.clip(np.asarray(l, dtype=np.float64), -inf, inf), dtype=dt)
        u = np.ascontiguousarray(np.clip(np.asarray(u, dtype=np.float64), -inf, inf), dtype=dt)
        if l.ndim != 2 or l.shape != u.shape or l.shape[1] != self.m:
            raise ValueError("l and u must have shape (nb, m)")
        nb = l.shape[0]
        qb = None
        if q is not None:
            qb = np.ascontiguousarray(q, dtype=dt)
            if qb.shape != (nb, self.n):
                raise ValueError("q must have shape (nb, n)")
        x = np.empty((nb, self.n), dtype=dt)
        y = np.empty((nb, self.m), dtype=dt)
        ints = [np.empty(nb, dtype=np.int32) for _ in range(4)]      # iters, status, cg_iters, rho_updates
        flts = [np.empty(nb, dtype=dt) for _ in range(3)]            # obj, prim_res, dual_res
        fp = C.POINTER(self._T.c_float)
        ip = C.POINTER(C.c_int)
        f = self._lib.osqp_b200_solve_batch
        f.restype = C.c_int
        f.argtypes = [C.c_void_p, C.c_int, fp, fp, fp, fp, fp, ip, ip, fp, fp, fp, ip, ip]
        P = lambda a: a.ctypes.data_as(fp)
        rc = f(C.cast(self._solver, C.c_void_p), nb, P(l), P(u), P(qb) if qb is not None else None, P(x), P(y),
               ints[0].ctypes.data_as(ip), ints[1].ctypes.data_as(ip), P(flts[0]), P(flts[1]), P(flts[2]),
               ints[2].ctypes.data_as(ip), ints[3].ctypes.data_as(ip))
        if rc != 0:
            raise OSQPError(rc, "osqp_b200_solve_batch")
        return SimpleNamespace(x=x, y=y, iter=ints[0], status_val=ints[1], cg_iters=ints[2], rho_updates=ints[3],
                               obj_val=flts[0], prim_res=flts[1], dual_res=flts[2])
