"""ctypes bindings of OSQP's PRIVATE algebra interface as exported by a libosqp build
(/root/reference/include/private/algebra_vector.h, algebra_matrix.h, lin_alg.h).  This is the
boundary the reference's own `lin_alg_tester` exercises (tests/lin_alg/lin_alg_tester.cpp); the
helpers below mirror the smart-pointer wrappers of tests/osqp_api.h:22-88."""
import ctypes as C

import numpy as np
import scipy.sparse as sp

vp, ci, cd = C.c_void_p, C.c_int, C.c_double
dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int)

SIGS = {
    "osqp_algebra_init_libs": (ci, [ci]), "osqp_algebra_free_libs": (None, []),
    "OSQPVectorf_new": (vp, [vp, ci]), "OSQPVectorf_malloc": (vp, [ci]), "OSQPVectorf_calloc": (vp, [ci]),
    "OSQPVectorf_copy_new": (vp, [vp]), "OSQPVectorf_free": (None, [vp]),
    "OSQPVectorf_view": (vp, [vp, ci, ci]), "OSQPVectorf_view_free": (None, [vp]),
    "OSQPVectorf_length": (ci, [vp]), "OSQPVectorf_copy": (None, [vp, vp]),
    "OSQPVectorf_from_raw": (None, [vp, vp]), "OSQPVectorf_to_raw": (None, [vp, vp]),
    "OSQPVectori_new": (vp, [vp, ci]), "OSQPVectori_malloc": (vp, [ci]), "OSQPVectori_calloc": (vp, [ci]),
    "OSQPVectori_free": (None, [vp]), "OSQPVectori_length": (ci, [vp]),
    "OSQPVectori_from_raw": (None, [vp, vp]), "OSQPVectori_to_raw": (None, [vp, vp]),
    "OSQPVectorf_is_eq": (ci, [vp, vp, cd]),
    "OSQPVectorf_set_scalar": (None, [vp, cd]),
    "OSQPVectorf_set_scalar_conditional": (None, [vp, vp, cd, cd, cd]),
    "OSQPVectorf_round_to_zero": (None, [vp, cd]), "OSQPVectorf_mult_scalar": (None, [vp, cd]),
    "OSQPVectorf_plus": (None, [vp, vp, vp]), "OSQPVectorf_minus": (None, [vp, vp, vp]),
    "OSQPVectorf_add_scaled": (None, [vp, cd, vp, cd, vp]),
    "OSQPVectorf_add_scaled3": (None, [vp, cd, vp, cd, vp, cd, vp]),
    "OSQPVectorf_norm_inf": (cd, [vp]), "OSQPVectorf_scaled_norm_inf": (cd, [vp, vp]),
    "OSQPVectorf_norm_inf_diff": (cd, [vp, vp]), "OSQPVectorf_norm_1": (cd, [vp]),
    "OSQPVectorf_norm_2": (cd, [vp]),
    "OSQPVectorf_dot_prod": (cd, [vp, vp]), "OSQPVectorf_dot_prod_signed": (cd, [vp, vp, ci]),
    "OSQPVectorf_ew_prod": (None, [vp, vp, vp]), "OSQPVectorf_all_leq": (ci, [vp, vp]),
    "OSQPVectorf_ew_bound_vec": (None, [vp, vp, vp, vp]),
    "OSQPVectorf_project_polar_reccone": (None, [vp, vp, vp, cd]),
    "OSQPVectorf_in_reccone": (ci, [vp, vp, vp, cd, cd]),
    "OSQPVectorf_ew_reciprocal": (None, [vp, vp]), "OSQPVectorf_ew_sqrt": (None, [vp]),
    "OSQPVectorf_ew_max_vec": (None, [vp, vp, vp]), "OSQPVectorf_ew_min_vec": (None, [vp, vp, vp]),
    "OSQPVectorf_ew_bounds_type": (ci, [vp, vp, vp, cd, cd]),
    "OSQPVectorf_set_scalar_if_lt": (None, [vp, vp, cd, cd]),
    "OSQPVectorf_set_scalar_if_gt": (None, [vp, vp, cd, cd]),
    "OSQPMatrix_new_from_csc": (vp, [vp, ci]), "OSQPMatrix_free": (None, [vp]),
    "OSQPMatrix_get_m": (ci, [vp]), "OSQPMatrix_get_n": (ci, [vp]), "OSQPMatrix_get_nz": (ci, [vp]),
    "OSQPMatrix_is_eq": (ci, [vp, vp, cd]),
    "OSQPMatrix_update_values": (None, [vp, vp, vp, ci]),
    "OSQPMatrix_mult_scalar": (None, [vp, cd]),
    "OSQPMatrix_lmult_diag": (None, [vp, vp]), "OSQPMatrix_rmult_diag": (None, [vp, vp]),
    "OSQPMatrix_Axpy": (None, [vp, vp, vp, cd, cd]), "OSQPMatrix_Atxpy": (None, [vp, vp, vp, cd, cd]),
    "OSQPMatrix_col_norm_inf": (None, [vp, vp]), "OSQPMatrix_row_norm_inf": (None, [vp, vp]),
    "OSQPMatrix_submatrix_byrows": (vp, [vp, vp]),
}


class Algebra:
    """Thin object layer over the private interface of one loaded libosqp."""

    def __init__(self, loaded, T):
        self.lib = loaded.lib
        self.T = T
        for name, (res, args) in SIGS.items():
            if hasattr(self.lib, name):
                fn = getattr(self.lib, name)
                fn.restype, fn.argtypes = res, args

    # -- vectors -----------------------------------------------------------------
    def vec(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        return self.lib.OSQPVectorf_new(a.ctypes.data, a.size)

    def veci(self, a):
        a = np.ascontiguousarray(a, dtype=np.int32)
        return self.lib.OSQPVectori_new(a.ctypes.data, a.size)

    def get(self, v):
        n = self.lib.OSQPVectorf_length(v)
        out = np.empty(n)
        if n:
            self.lib.OSQPVectorf_to_raw(out.ctypes.data, v)
        return out

    def geti(self, v):
        n = self.lib.OSQPVectori_length(v)
        out = np.empty(n, dtype=np.int32)
        if n:
            self.lib.OSQPVectori_to_raw(out.ctypes.data, v)
        return out

    # -- matrices ----------------------------------------------------------------
    def mat(self, M, triu=False):
        M = sp.csc_matrix(M)
        M.sort_indices()
        keep = []
        p = np.ascontiguousarray(M.indptr, dtype=np.int32)
        i = np.ascontiguousarray(M.indices, dtype=np.int32)
        x = np.ascontiguousarray(M.data, dtype=np.float64)
        s = self.T.OSQPCscMatrix()
        s.m, s.n = M.shape
        s.p, s.i, s.x = p.ctypes.data_as(ip), i.ctypes.data_as(ip), x.ctypes.data_as(dp)
        s.nzmax, s.nz, s.owned = int(M.nnz), -1, 0
        h = self.lib.OSQPMatrix_new_from_csc(C.addressof(s), 1 if triu else 0)
        assert h, "OSQPMatrix_new_from_csc failed"
        return h

    def axpy(self, M, x, y, alpha, beta, transpose=False):
        vx, vy = self.vec(x), self.vec(y)
        (self.lib.OSQPMatrix_Atxpy if transpose else self.lib.OSQPMatrix_Axpy)(M, vx, vy, alpha, beta)
        out = self.get(vy)
        self.lib.OSQPVectorf_free(vx)
        self.lib.OSQPVectorf_free(vy)
        return out
