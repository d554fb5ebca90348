"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/osqp_b200.h declares, libosqp_b200 exports the unchanged public OSQP API plus the 60
private-interface symbols the core links against, and the ctypes mirrors
match the compiled struct layouts.  No compute calls (no GPU here)."""
import ctypes as C
import subprocess

import pytest

from osqp_b200 import _capi
from osqp_b200._lib import declared_symbols, lib_paths, load_kernels, load_library

# linker-verified backend symbol set
BACKEND_SYMBOLS = (
    ["osqp_algebra_" + s for s in ("default_linsys", "device_name", "free_libs", "init_libs",
                                   "init_linsys_solver", "linsys_supported", "name")]
    + ["OSQPMatrix_" + s for s in ("Atxpy", "Axpy", "col_norm_inf", "free", "get_m", "get_nz",
                                   "lmult_diag", "mult_scalar", "new_from_csc", "rmult_diag",
                                   "row_norm_inf", "submatrix_byrows", "update_values")]
    + ["OSQPVectorf_" + s for s in (
        "add_scaled", "add_scaled3", "all_leq", "calloc", "copy", "copy_new", "dot_prod",
        "dot_prod_signed", "ew_bound_vec", "ew_bounds_type", "ew_max_vec", "ew_prod",
        "ew_reciprocal", "ew_sqrt", "free", "from_raw", "in_reccone", "length", "malloc", "minus",
        "mult_scalar", "new", "norm_1", "norm_inf", "plus", "project_polar_reccone",
        "round_to_zero", "scaled_norm_inf", "set_scalar", "set_scalar_conditional",
        "set_scalar_if_gt", "set_scalar_if_lt", "to_raw", "view", "view_free")]
    + ["OSQPVectori_" + s for s in ("calloc", "free", "from_raw", "malloc", "to_raw")])
PUBLIC_API = ["osqp_setup", "osqp_solve", "osqp_cleanup", "osqp_update_data_vec",
              "osqp_update_data_mat", "osqp_update_rho", "osqp_update_settings", "osqp_warm_start",
              "osqp_cold_start", "osqp_get_solution", "osqp_set_default_settings", "osqp_version"]


def test_header_declares_the_whole_cabi():
    syms = declared_symbols()
    assert len(syms) >= 60
    for family in ("b200_vec_", "b200_csr_", "b200_pcg_", "b200_admm_"):
        assert any(s.startswith(family) for s in syms)


@pytest.mark.parametrize("precision", ["f64"])
def test_kernel_library_exports_every_declared_symbol(precision):
    k = load_kernels(precision)
    missing = [s for s in declared_symbols() if not hasattr(k, s)]
    assert not missing, f"declared in include/osqp_b200.h but not exported: {missing}"


def test_backend_symbol_set_is_complete():
    assert len(BACKEND_SYMBOLS) == 60
    L = load_library("f64")
    missing = [s for s in BACKEND_SYMBOLS + PUBLIC_API if not hasattr(L.lib, s)]
    assert not missing, f"libosqp_b200 lacks: {missing}"


def test_no_vendor_sparse_or_blas_libraries_linked():
    """north_star: no cuSPARSE, cuBLAS or thrust."""
    for path in lib_paths("f64"):
        out = subprocess.run(["ldd", str(path)], capture_output=True, text=True).stdout.lower()
        for banned in ("cusparse", "cublas", "cusolver"):
            assert banned not in out, f"{path} links {banned}"


def test_ctypes_mirror_matches_compiled_layout():
    L = load_library("f64")
    T = _capi.TYPES_F64
    assert L.lib.osqp_b200_sizeof(0) == C.sizeof(T.OSQPSettings)
    assert L.lib.osqp_b200_sizeof(1) == C.sizeof(T.OSQPInfo)
    assert L.lib.osqp_b200_sizeof(2) == C.sizeof(T.OSQPCscMatrix)
    assert L.lib.osqp_b200_sizeof(3) == 8 and L.lib.osqp_b200_sizeof(4) == 4


def test_default_settings_are_the_reference_defaults():
    """osqp_set_default_settings (src/osqp_api.c:283-327); a new backend macro gets the CPU
    values of the backend-conditional defaults (osqp_api_constants.h:111-178)."""
    from osqp_b200.interface import OSQP
    s = OSQP(load_library("f64")).default_settings()
    assert s.linsys_solver == _capi.OSQP_INDIRECT_SOLVER   # only solver the backend offers
    assert (s.rho, s.sigma, s.alpha) == (0.1, 1e-6, 1.6)
    assert (s.cg_max_iter, s.cg_tol_reduction, s.cg_tol_fraction) == (20, 10, 0.15)
    assert s.cg_precond == _capi.OSQP_DIAGONAL_PRECONDITIONER
    assert (s.rho_is_vec, s.check_termination, s.adaptive_rho_tolerance) == (1, 25, 5.0)
    assert (s.eps_abs, s.eps_rel, s.max_iter, s.scaling) == (1e-3, 1e-3, 4000, 10)


def test_product_path_fails_loudly_without_gpu():
    """There is no CPU fallback: without a device osqp_setup must return
    OSQP_ALGEBRA_LOAD_ERROR (7), never a silent CPU solve."""
    import numpy as np
    import scipy.sparse as sp
    from osqp_b200.devmem import kernels
    from osqp_b200.interface import OSQP, OSQPError
    k = kernels("f64")
    if k.b200_init(0) == 0:
        k.b200_shutdown()
        pytest.skip("a GPU is present")
    with pytest.raises(OSQPError) as e:
        OSQP(load_library("f64")).setup(sp.eye(2, format="csc"), np.ones(2), sp.eye(2, format="csc"),
                                        -np.ones(2), np.ones(2))
    assert e.value.code == 7
