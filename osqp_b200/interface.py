"""Host-side mirror of OSQP's user-facing solver object.

`OSQP` wraps the *unchanged* public C API (`osqp_setup` / `osqp_solve` /
`osqp_update_*`, /root/reference/include/public/osqp_api_functions.h:267-454) of a
libosqp build.  The product build links the reference core against the B200
algebra backend (`algebra/b200`), so every vector/matrix/linear-solve call the
core makes lands in hand-written sm_100a kernels.  The class is library-agnostic:
tests hand it the CPU oracle library to get reference answers through the very
same calls (method names follow the reference's Python wrapper as used in
/root/reference/docs/examples/*.rst: setup / solve / update / warm_start /
update_settings).
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import scipy.sparse as sp

from . import _capi


class OSQPError(RuntimeError):
    def __init__(self, code, where):
        super().__init__(f"{where} failed with OSQP error code {code}")
        self.code = code


class LoadedLibrary:
    """A libosqp shared object plus the ctypes struct set matching its OSQPFloat."""

    def __init__(self, path, dtype=np.float64):
        self.path = str(path)
        self.dtype = np.dtype(dtype)
        self.T = _capi.TYPES_F64 if self.dtype == np.float64 else _capi.TYPES_F32
        self.lib = _capi.bind_public_api(C.CDLL(self.path, mode=C.RTLD_LOCAL), self.T)


def _csc_struct(T, M, dtype, keep):
    """Build an OSQPCscMatrix view over a scipy CSC matrix (no copies of ours
    outlive `keep`)."""
    if not sp.isspmatrix_csc(M):      # re-wrapping a CSC matrix would drop its cached sortedness flag
        M = sp.csc_matrix(M)
    M.sort_indices()
    p = np.ascontiguousarray(M.indptr, dtype=np.int32)
    i = np.ascontiguousarray(M.indices, dtype=np.int32)
    x = np.ascontiguousarray(M.data, dtype=dtype)
    keep.extend([p, i, x])
    s = T.OSQPCscMatrix()
    s.m, s.n = M.shape
    s.p = p.ctypes.data_as(C.POINTER(C.c_int))
    s.i = i.ctypes.data_as(C.POINTER(C.c_int))
    s.x = x.ctypes.data_as(C.POINTER(T.c_float))
    s.nzmax = int(M.nnz)
    s.nz = -1
    s.owned = 0
    return s


def _upper_triangle(P):
    """Upper triangle of P as CSC; P itself when it already is one (sp.triu goes through COO:
    24 ms for the 1e6-entry P of the Lasso workload, inside the end-to-end timed region)."""
    P = P if sp.isspmatrix_csc(P) else sp.csc_matrix(P)
    if P.nnz == 0:
        return P
    if P.has_sorted_indices:
        # sorted columns: the last entry of a column is its largest row index -- O(n), not O(nnz)
        ne = np.nonzero(np.diff(P.indptr))[0]
        if bool((P.indices[P.indptr[ne + 1] - 1] <= ne).all()):
            return P
    else:
        cols = np.repeat(np.arange(P.shape[1], dtype=P.indices.dtype), np.diff(P.indptr))
        if bool((P.indices <= cols).all()):
            return P
    return sp.triu(P, format="csc")


class OSQP:
    """Solver object: `setup`, `solve`, `update`, `warm_start`, `update_settings`."""

    def __init__(self, library):
        self._L = library
        self._lib = library.lib
        self._T = library.T
        self._solver = None
        self._keep = []
        self.n = self.m = 0

    # -- helpers ---------------------------------------------------------------
    def _fp(self, a):
        if a is None:
            return None
        a = np.ascontiguousarray(a, dtype=self._L.dtype)
        self._keep.append(a)
        return a.ctypes.data_as(C.POINTER(self._T.c_float))

    def default_settings(self):
        s = self._T.OSQPSettings()
        self._lib.osqp_set_default_settings(C.byref(s))
        return s

    @staticmethod
    def _apply(settings, kw):
        for k, v in kw.items():
            if not hasattr(settings, k):
                raise TypeError(f"unknown OSQP setting {k!r}")
            setattr(settings, k, v)

    # -- API -------------------------------------------------------------------
    def setup(self, P, q, A, l, u, **settings):
        T = self._T
        P = _upper_triangle(P)
        A = A if sp.isspmatrix_csc(A) else sp.csc_matrix(A)
        self.n = P.shape[0]
        self.m = A.shape[0]
        if A.shape[1] != self.n:
            raise ValueError("A must have n columns")
        inf = _capi.OSQP_INFTY
        l = np.clip(np.asarray(l, dtype=np.float64), -inf, inf)
        u = np.clip(np.asarray(u, dtype=np.float64), -inf, inf)
        st = self.default_settings()
        st.verbose = 0
        self._apply(st, settings)
        keep = []
        Ps = _csc_struct(T, P, self._L.dtype, keep)
        As = _csc_struct(T, A, self._L.dtype, keep)
        self._nnzP, self._nnzA = P.nnz, A.nnz
        solver = C.POINTER(T.OSQPSolver)()
        self._keep = keep
        rc = self._lib.osqp_setup(C.byref(solver), C.byref(Ps), self._fp(q),
                                  C.byref(As), self._fp(l), self._fp(u),
                                  self.m, self.n, C.byref(st))
        self._keep = []
        if rc != 0:
            raise OSQPError(rc, "osqp_setup")
        self._solver = solver
        return self

    @property
    def settings(self):
        return self._solver.contents.settings.contents

    @property
    def info(self):
        return self._solver.contents.info.contents

    def solve(self):
        rc = self._lib.osqp_solve(self._solver)
        if rc != 0:
            raise OSQPError(rc, "osqp_solve")
        return self.results()

    def results(self):
        sol = self._solver.contents.solution.contents
        info = self.info
        dt = self._L.dtype

        def arr(ptr, n):
            if n == 0 or not ptr:
                return np.zeros(0, dtype=dt)
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True)

        def cert(ptr, n):
            # store_solution (auxil.c:627-631) fills a certificate that does not apply with OSQP_NAN: copying
            # (n + m) marker values per solve costs as much as copying the solution itself -- None instead
            if n == 0 or not ptr or (ptr[0] == _capi.OSQP_NAN and ptr[n - 1] == _capi.OSQP_NAN):
                return None
            return arr(ptr, n)

        inf = SimpleNamespace(**{k: getattr(info, k) for k, _ in info._fields_})
        inf.status = info.status.decode()
        return SimpleNamespace(x=arr(sol.x, self.n), y=arr(sol.y, self.m),
                               prim_inf_cert=cert(sol.prim_inf_cert, self.m),
                               dual_inf_cert=cert(sol.dual_inf_cert, self.n),
                               info=inf)

    def update(self, q=None, l=None, u=None, Px=None, Px_idx=None, Ax=None, Ax_idx=None):
        inf = _capi.OSQP_INFTY
        if l is not None:
            l = np.clip(np.asarray(l, dtype=np.float64), -inf, inf)
        if u is not None:
            u = np.clip(np.asarray(u, dtype=np.float64), -inf, inf)
        if q is not None or l is not None or u is not None:
            rc = self._lib.osqp_update_data_vec(self._solver, self._fp(q), self._fp(l), self._fp(u))
            self._keep = []
            if rc != 0:
                raise OSQPError(rc, "osqp_update_data_vec")
        if Px is not None or Ax is not None:
            def ip(a):
                if a is None:
                    return None
                a = np.ascontiguousarray(a, dtype=np.int32)
                self._keep.append(a)
                return a.ctypes.data_as(C.POINTER(C.c_int))
            nP = 0 if Px is None else len(Px)
            nA = 0 if Ax is None else len(Ax)
            rc = self._lib.osqp_update_data_mat(self._solver, self._fp(Px), ip(Px_idx), nP,
                                                self._fp(Ax), ip(Ax_idx), nA)
            self._keep = []
            if rc != 0:
                raise OSQPError(rc, "osqp_update_data_mat")

    def warm_start(self, x=None, y=None):
        rc = self._lib.osqp_warm_start(self._solver, self._fp(x), self._fp(y))
        self._keep = []
        if rc != 0:
            raise OSQPError(rc, "osqp_warm_start")

    def cold_start(self):
        self._lib.osqp_cold_start(self._solver)

    def update_settings(self, **kw):
        T = self._T
        st = T.OSQPSettings()
        C.memmove(C.byref(st), C.byref(self.settings), C.sizeof(st))
        rho = kw.pop("rho", None)
        self._apply(st, kw)
        rc = self._lib.osqp_update_settings(self._solver, C.byref(st))
        if rc != 0:
            raise OSQPError(rc, "osqp_update_settings")
        if rho is not None:
            rc = self._lib.osqp_update_rho(self._solver, rho)
            if rc != 0:
                raise OSQPError(rc, "osqp_update_rho")

    def cleanup(self):
        if self._solver is not None:
            self._lib.osqp_cleanup(self._solver)
            self._solver = None

    def __del__(self):
        try:
            self.cleanup()
        except Exception:
            pass
