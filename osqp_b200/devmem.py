"""Minimal device-array helper over the C-ABI (b200_malloc / b200_copy_in / b200_copy_out).
Used by tests and bench.py to drive individual kernels; the solver itself never needs it."""
import ctypes as C

import numpy as np

from ._lib import load_kernels

_bound = {}


def kernels(precision="f64"):
    """Kernel library with argtypes/restypes declared from include/osqp_b200.h."""
    if precision in _bound:
        return _bound[precision]
    k = load_kernels(precision)
    F = C.c_double if precision == "f64" else C.c_float
    vp, ip, i, sz = C.c_void_p, C.c_void_p, C.c_int, C.c_size_t
    sig = {
        "b200_init": (i, [i]), "b200_shutdown": (None, []), "b200_sync": (None, []),
        "b200_device_name": (i, [C.c_char_p, i]), "b200_sm_count": (i, []),
        "b200_last_error": (i, []), "b200_launch_count": (C.c_ulonglong, []),
        "b200_stream_handle": (vp, []),
        "b200_event_create": (vp, []), "b200_event_destroy": (None, [vp]),
        "b200_event_record": (None, [vp]), "b200_event_elapsed_ms": (C.c_float, [vp, vp]),
        "b200_malloc": (vp, [sz]), "b200_calloc": (vp, [sz]), "b200_free": (None, [vp]),
        "b200_copy_in": (i, [vp, vp, sz]), "b200_copy_out": (i, [vp, vp, sz]),
        "b200_ptr_is_device": (i, [vp]),
        "b200_vec_set_scalar": (None, [vp, F, i]),
        "b200_vec_set_scalar_cond": (None, [vp, ip, F, F, F, i]),
        "b200_vec_round_to_zero": (None, [vp, F, i]),
        "b200_vec_mult_scalar": (None, [vp, F, i]),
        "b200_vec_add_scaled": (None, [vp, F, vp, F, vp, i]),
        "b200_vec_add_scaled3": (None, [vp, F, vp, F, vp, F, vp, i]),
        "b200_vec_ew_prod": (None, [vp, vp, vp, i]),
        "b200_vec_ew_bound": (None, [vp, vp, vp, vp, i]),
        "b200_vec_project_polar_reccone": (None, [vp, vp, vp, F, i]),
        "b200_vec_ew_reciprocal": (None, [vp, vp, i]),
        "b200_vec_ew_sqrt": (None, [vp, i]),
        "b200_vec_ew_max": (None, [vp, vp, vp, i]),
        "b200_vec_ew_min": (None, [vp, vp, vp, i]),
        "b200_vec_set_scalar_if_lt": (None, [vp, vp, F, F, i]),
        "b200_vec_set_scalar_if_gt": (None, [vp, vp, F, F, i]),
        "b200_vec_scatter": (None, [vp, vp, ip, i]),
        "b200_vec_gather": (None, [vp, vp, ip, i]),
        "b200_vec_norm_inf": (F, [vp, i]),
        "b200_vec_scaled_norm_inf": (F, [vp, vp, i]),
        "b200_vec_norm_inf_diff": (F, [vp, vp, i]),
        "b200_vec_norm_1": (F, [vp, i]), "b200_vec_norm_2": (F, [vp, i]),
        "b200_vec_dot": (F, [vp, vp, i]),
        "b200_vec_dot_signed": (F, [vp, vp, i, i]),
        "b200_vec_all_leq": (i, [vp, vp, i]),
        "b200_vec_in_reccone": (i, [vp, vp, vp, F, F, i]),
        "b200_vec_is_eq": (i, [vp, vp, F, i]),
        "b200_veci_is_eq": (i, [ip, ip, i]),
        "b200_vec_bounds_type": (i, [ip, vp, vp, F, F, i]),
        "b200_csr_create": (vp, [i, i, i, ip, ip, vp]),
        "b200_csr_destroy": (None, [vp]),
        "b200_csr_transpose": (vp, [vp, C.POINTER(C.c_void_p)]),
        "b200_veci_gather": (None, [ip, ip, ip, i]),
        "b200_csr_symmetric_from_triu": (vp, [i, ip, ip, vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
        "b200_vec_scatter_nonneg": (None, [vp, vp, ip, i]),
        "b200_csr_nrows": (i, [vp]), "b200_csr_ncols": (i, [vp]), "b200_csr_nnz": (i, [vp]),
        "b200_csr_values": (vp, [vp]),
        "b200_csr_download": (i, [vp, ip, ip, vp]),
        "b200_csr_spmv": (None, [vp, vp, vp, F, F]),
        "b200_csr_scale": (None, [vp, F]),
        "b200_csr_scale_rows": (None, [vp, vp]), "b200_csr_scale_cols": (None, [vp, vp]),
        "b200_csr_row_absmax": (None, [vp, vp]),
        "b200_csr_row_absmax_lower": (None, [vp, vp]),
        "b200_csr_row_wsumsq": (None, [vp, vp, F, vp]),
        "b200_csr_diag": (None, [vp, vp]),
        "b200_csr_is_eq": (i, [vp, vp, F]),
        "b200_pcg_create": (vp, [vp, vp, vp, i, i]),
        "b200_pcg_destroy": (None, [vp]),
        "b200_pcg_configure": (None, [vp, F, F, vp, i, i]),
        "b200_pcg_refresh_matrices": (None, [vp]), "b200_pcg_refresh_precond": (None, [vp]),
        "b200_pcg_warm_start": (None, [vp, vp]),
        "b200_pcg_solve": (i, [vp, vp, i, C.c_double, C.c_double, i, C.c_double, i]),
        "b200_pcg_stats": (None, [vp, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong),
                                  C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
        "b200_admm_compute_rhs": (None, [vp, vp, vp, vp, vp, vp, vp, F, F, i, i]),
        "b200_admm_update_xzy": (None, [vp] * 13 + [F, F, F, i, i]),
        "b200_admm_update_xzy_carry": (None, [vp] * 13 + [F, F, F, i, i, vp]),
        "b200_admm_residuals": (None, [vp] * 11 + [F, F, i, i, vp]),
        "b200_admm_infeas_scalars": (None, [vp] * 7 + [F, i, i, i, i, vp]),
        "b200_epoch": (C.c_ulonglong, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(k, name)
        fn.restype = res
        fn.argtypes = args
    k._sig = sig
    k._ftype = np.float64 if precision == "f64" else np.float32
    _bound[precision] = k
    return k


class DeviceArray:
    """Owning handle of a device buffer holding a numpy-typed 1-D array."""

    def __init__(self, k, host=None, n=None, dtype=None):
        self.k = k
        if host is not None:
            host = np.ascontiguousarray(host)
            dtype, n = host.dtype, host.size
        self.dtype = np.dtype(dtype if dtype is not None else k._ftype)
        self.n = int(n)
        self.ptr = k.b200_malloc(max(self.n, 1) * self.dtype.itemsize)
        if not self.ptr:
            raise MemoryError("b200_malloc failed")
        if host is not None and self.n:
            k.b200_copy_in(self.ptr, host.ctypes.data, host.nbytes)

    def get(self):
        out = np.empty(self.n, dtype=self.dtype)
        if self.n:
            self.k.b200_copy_out(out.ctypes.data, self.ptr, out.nbytes)
        return out

    def offset(self, elems):
        return C.c_void_p(self.ptr + elems * self.dtype.itemsize)

    def free(self):
        if self.ptr:
            self.k.b200_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def csr_to_device(k, M):
    """Upload a scipy matrix as a b200_csr handle (caller destroys it)."""
    import scipy.sparse as sp
    M = sp.csr_matrix(M)
    M.sort_indices()
    rp = np.ascontiguousarray(M.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(M.indices, dtype=np.int32)
    vx = np.ascontiguousarray(M.data, dtype=k._ftype)
    h = k.b200_csr_create(M.shape[0], M.shape[1], int(M.nnz), rp.ctypes.data, ci.ctypes.data,
                          vx.ctypes.data)
    if not h:
        raise MemoryError("b200_csr_create failed")
    return h
