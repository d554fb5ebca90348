"""GPU parity tests at the drop-in boundary: OSQP's private algebra interface as implemented by
algebra/b200 (plain C) over the sm_100a kernels, checked against the reference's OWN golden
vectors (tests/lin_alg/generate_problem.py, captured in tests/golden/lin_alg.npz).  The cases
mirror tests/lin_alg/testcases/test_vector_*.cpp, test_mat_vec.cpp and test_matrix.cpp, and are
also run against the CPU oracle so that both backends are held to the same answers."""
import numpy as np
import pytest
import scipy.sparse as sp

from algebra_api import Algebra
from conftest import TESTS_TOL, load_golden
from osqp_b200 import _capi

pytestmark = pytest.mark.gpu
G = load_golden("lin_alg")
TOL = TESTS_TOL


@pytest.fixture(scope="module", params=["b200", "oracle"])
def alg(request, b200_lib, oracle_lib):
    L = b200_lib if request.param == "b200" else oracle_lib
    a = Algebra(L, _capi.TYPES_F64)
    assert a.lib.osqp_algebra_init_libs(0) == 0
    yield a
    a.lib.osqp_algebra_free_libs()


def v(name):
    return np.asarray(G["test_vec_ops_" + name], dtype=float)


# ---------------------------------------------------------------- test_vector_math_ops.cpp
def test_plus_minus_including_in_place(alg):
    a, b = alg.vec(v("v1")), alg.vec(v("v2"))
    r = alg.vec(np.zeros(10))
    alg.lib.OSQPVectorf_plus(r, a, b)
    assert np.abs(alg.get(r) - v("add")).max() < TOL
    alg.lib.OSQPVectorf_minus(r, a, b)
    assert np.abs(alg.get(r) - v("sub")).max() < TOL
    alg.lib.OSQPVectorf_plus(a, a, b)            # x == a
    assert np.abs(alg.get(a) - v("add")).max() < TOL
    alg.lib.OSQPVectorf_minus(a, a, b)
    assert np.abs(alg.get(a) - v("v1")).max() < TOL


def test_add_scaled_and_add_scaled3(alg):
    a, b, c = alg.vec(v("v1")), alg.vec(v("v2")), alg.vec(v("v3"))
    r = alg.vec(np.zeros(10))
    s1, s2, s3 = (float(G["test_vec_ops_sc%d" % i]) for i in (1, 2, 3))
    alg.lib.OSQPVectorf_add_scaled(r, s1, a, s2, b)
    assert np.abs(alg.get(r) - v("add_scaled")).max() < TOL
    alg.lib.OSQPVectorf_add_scaled3(r, s1, a, s2, b, s3, c)
    assert np.abs(alg.get(r) - v("add_scaled3")).max() < TOL
    a2 = alg.vec(v("v1"))
    alg.lib.OSQPVectorf_add_scaled(a2, 1.0, a2, s2, b)          # accumulate form
    assert np.abs(alg.get(a2) - v("add_scaled_inc")).max() < TOL
    a3 = alg.vec(v("v1"))
    alg.lib.OSQPVectorf_add_scaled3(a3, 1.0, a3, s2, b, s3, c)
    assert np.abs(alg.get(a3) - v("add_scaled3_inc")).max() < TOL


def test_scalar_and_elementwise_products(alg):
    a, b = alg.vec(v("v1")), alg.vec(v("v2"))
    r = alg.vec(np.zeros(10))
    alg.lib.OSQPVectorf_ew_prod(r, a, b)
    assert np.abs(alg.get(r) - v("ew_prod")).max() < TOL
    alg.lib.OSQPVectorf_ew_prod(a, a, b)                          # c == a
    assert np.abs(alg.get(a) - v("ew_prod")).max() < TOL
    a = alg.vec(v("v1"))
    alg.lib.OSQPVectorf_mult_scalar(a, float(G["test_vec_ops_sc1"]))
    assert np.abs(alg.get(a) - v("sca_prod")).max() < TOL


def test_sqrt_reciprocal_max_min(alg):
    a = alg.vec(v("shift_v1"))
    alg.lib.OSQPVectorf_ew_sqrt(a)
    assert np.abs(alg.get(a) - v("ew_sqrt")).max() < TOL
    a, r = alg.vec(v("v1")), alg.vec(np.zeros(10))
    alg.lib.OSQPVectorf_ew_reciprocal(r, a)
    assert np.abs(alg.get(r) - v("ew_reciprocal")).max() < TOL * np.abs(v("ew_reciprocal")).max()
    b = alg.vec(v("v2"))
    alg.lib.OSQPVectorf_ew_max_vec(r, a, b)
    assert np.abs(alg.get(r) - v("ew_max_vec")).max() < TOL
    alg.lib.OSQPVectorf_ew_min_vec(r, a, b)
    assert np.abs(alg.get(r) - v("ew_min_vec")).max() < TOL


def test_ew_bound_vec(alg):
    # x = min(max(z, l), u) with (z, l, u) = (v1, v3, v2) as in generate_problem.py:71
    z, l, u = alg.vec(v("v1")), alg.vec(v("v3")), alg.vec(v("v2"))
    r = alg.vec(np.zeros(10))
    alg.lib.OSQPVectorf_ew_bound_vec(r, z, l, u)
    assert np.abs(alg.get(r) - v("ew_bound_vec")).max() < TOL
    alg.lib.OSQPVectorf_ew_bound_vec(z, z, l, u)                 # in place
    assert np.abs(alg.get(z) - v("ew_bound_vec")).max() < TOL
    # infinite bounds leave z alone
    z = alg.vec(v("v1"))
    lo, hi = alg.vec(-1e30 * np.ones(10)), alg.vec(1e30 * np.ones(10))
    alg.lib.OSQPVectorf_ew_bound_vec(r, z, lo, hi)
    assert np.abs(alg.get(r) - v("v1")).max() == 0.0


def test_norms_and_dots(alg):
    a, b = alg.vec(v("v1")), alg.vec(v("v2"))
    L = alg.lib
    assert abs(L.OSQPVectorf_norm_inf(a) - G["test_vec_ops_norm_inf"]) < TOL
    assert abs(L.OSQPVectorf_norm_1(a) - G["test_vec_ops_norm_1"]) < TOL
    assert abs(L.OSQPVectorf_scaled_norm_inf(a, b) - G["test_vec_ops_norm_inf_scaled"]) < TOL
    assert abs(L.OSQPVectorf_norm_inf_diff(a, b) - G["test_vec_ops_norm_inf_diff"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod(a, b) - G["test_vec_ops_vec_dot"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod(a, a) - G["test_vec_ops_vec_dot_v1"]) < TOL
    e = alg.vec(np.zeros(0))
    assert L.OSQPVectorf_norm_inf(e) == 0.0 and L.OSQPVectorf_dot_prod(e, e) == 0.0


def test_dot_prod_signed(alg):
    a, b = alg.vec(v("v1")), alg.vec(v("v2"))
    L = alg.lib
    assert abs(L.OSQPVectorf_dot_prod_signed(a, b, +1) - G["test_vec_ops_vec_dot_pos"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod_signed(a, b, -1) - G["test_vec_ops_vec_dot_neg"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod_signed(b, a, +1) - G["test_vec_ops_vec_dot_pos_flip"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod_signed(b, a, -1) - G["test_vec_ops_vec_dot_neg_flip"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod_signed(a, a, +1) - G["test_vec_ops_vec_dot_pos_v1"]) < TOL
    assert abs(L.OSQPVectorf_dot_prod_signed(a, b, 0) - G["test_vec_ops_vec_dot"]) < TOL   # fallback


def test_round_to_zero(alg):
    x = np.array([1e-16, -1e-16, 1e-3, -2.0, 0.0, 1e-15])
    a = alg.vec(x)
    alg.lib.OSQPVectorf_round_to_zero(a, 1e-15)                  # strict: |a| < tol
    assert (alg.get(a) == np.array([0, 0, 1e-3, -2.0, 0, 1e-15])).all()


# ---------------------------------------------------------------- test_vector_creation.cpp
def test_creation_copy_views_and_set_scalar(alg):
    L = alg.lib
    a = alg.vec(v("v1"))
    assert L.OSQPVectorf_length(a) == 10
    c = L.OSQPVectorf_copy_new(a)
    assert (alg.get(c) == v("v1")).all()
    z = L.OSQPVectorf_calloc(10)
    assert (alg.get(z) == 0).all()
    L.OSQPVectorf_set_scalar(z, float(G["test_vec_ops_sc1"]))
    assert (alg.get(z) == v("same")).all()
    view = L.OSQPVectorf_view(a, 3, 4)
    assert L.OSQPVectorf_length(view) == 4 and (alg.get(view) == v("v1")[3:7]).all()
    L.OSQPVectorf_set_scalar(view, 7.0)                  # writes through to the parent
    assert (alg.get(a)[3:7] == 7.0).all() and alg.get(a)[2] == v("v1")[2]
    L.OSQPVectorf_view_free(view)
    for h in (a, c, z):
        L.OSQPVectorf_free(h)


def test_set_scalar_conditional_and_if_lt_gt(alg):
    L = alg.lib
    s1, s2, s3 = (float(G["test_vec_ops_sc%d" % i]) for i in (1, 2, 3))
    r, cond = alg.vec(np.zeros(10)), alg.veci(G["test_vec_ops_sca_cond"])
    L.OSQPVectorf_set_scalar_conditional(r, cond, s1, s2, s3)
    assert (alg.get(r) == v("sca_cond_res")).all()
    a = alg.vec(v("v1"))
    L.OSQPVectorf_set_scalar_if_lt(r, a, s1, s2)
    assert (alg.get(r) == v("sca_lt")).all()
    L.OSQPVectorf_set_scalar_if_gt(r, a, s1, s2)
    assert (alg.get(r) == v("sca_gt")).all()


# ---------------------------------------------------------------- test_vector_comparisons.cpp
def test_is_eq_and_all_leq(alg):
    L = alg.lib
    a, b, c = alg.vec(v("v1")), alg.vec(v("v1")), alg.vec(v("v1") + 1.0)
    assert L.OSQPVectorf_is_eq(a, b, TOL) == 1 and L.OSQPVectorf_is_eq(a, c, TOL) == 0
    assert L.OSQPVectorf_all_leq(a, c) == 1 and L.OSQPVectorf_all_leq(c, a) == 0
    assert L.OSQPVectorf_all_leq(a, b) == 1


def test_bounds_type_reccone_projection(alg):
    """auxil.c:85-89 / builtin/vector.c:683-733,888-922"""
    L = alg.lib
    inf = 1e30 * 1e-4
    l = np.array([-1e30, -1e30, 0.0, 1.0, 2.0])
    u = np.array([1e30, 3.0, 1e30, 1.0, 2.00001])
    vl, vu = alg.vec(l), alg.vec(u)
    t = alg.veci(np.zeros(5, dtype=np.int32))
    changed = L.OSQPVectorf_ew_bounds_type(t, vl, vu, 1e-4, inf)
    assert changed == 1 and (alg.geti(t) == np.array([-1, 0, 0, 1, 1])).all()
    assert L.OSQPVectorf_ew_bounds_type(t, vl, vu, 1e-4, inf) == 0     # nothing changed now
    y = alg.vec(np.array([1.0, 2.0, -3.0, 4.0, -5.0]))
    L.OSQPVectorf_project_polar_reccone(y, vl, vu, inf)
    assert (alg.get(y) == np.array([0.0, 2.0, -3.0, 4.0, -5.0])).all()
    y = alg.vec(np.array([1.0, -2.0, 3.0, 4.0, -5.0]))
    L.OSQPVectorf_project_polar_reccone(y, vl, vu, inf)
    assert (alg.get(y) == np.array([0.0, 0.0, 0.0, 4.0, -5.0])).all()
    yy = alg.vec(np.array([5.0, -1.0, 1.0, 0.0, 0.0]))
    assert L.OSQPVectorf_in_reccone(yy, vl, vu, inf, 1e-6) == 1
    yy = alg.vec(np.array([5.0, 1.0, 1.0, 0.0, 0.0]))          # u[1] finite and y[1] > tol
    assert L.OSQPVectorf_in_reccone(yy, vl, vu, inf, 1e-6) == 0


# ---------------------------------------------------------------- test_mat_vec.cpp
def test_mat_vec_products(alg):
    A, Pu = G["test_mat_vec_A"], G["test_mat_vec_Pu"]
    x, y = G["test_mat_vec_x"], G["test_mat_vec_y"]
    hA, hP = alg.mat(A), alg.mat(Pu, triu=True)
    assert np.abs(alg.axpy(hA, x, np.full(5, np.nan), 1.0, 0.0) - G["test_mat_vec_Ax"]).max() < TOL
    assert np.abs(alg.axpy(hA, x, y, 1.0, 1.0) - G["test_mat_vec_Ax_cum"]).max() < TOL
    assert np.abs(alg.axpy(hA, y, np.zeros(4), 1.0, 0.0, True) - G["test_mat_vec_ATy"]).max() < TOL
    assert np.abs(alg.axpy(hA, y, x, 1.0, 1.0, True) - G["test_mat_vec_ATy_cum"]).max() < TOL
    # symmetric product from the upper triangle (csc_Axpy_sym_triu)
    assert np.abs(alg.axpy(hP, x, np.zeros(4), 1.0, 0.0) - G["test_mat_vec_Px"]).max() < TOL
    assert np.abs(alg.axpy(hP, x, x, 1.0, 1.0) - G["test_mat_vec_Px_cum"]).max() < TOL
    assert np.abs(alg.axpy(hP, x, x, -2.0, 0.5) - (-2 * G["test_mat_vec_Px"] + 0.5 * x)).max() < TOL
    assert alg.lib.OSQPMatrix_get_nz(hP) == Pu.nnz        # triu count (osqp_api.c:1333-1347)
    alg.lib.OSQPMatrix_free(hA)
    alg.lib.OSQPMatrix_free(hP)


def test_mat_vec_empty_cases(alg):
    """tests/lin_alg/test_mat_vec.cpp:71-149"""
    two = np.asarray(G["test_vec_mat_empty"], dtype=float)
    h = alg.mat(G["test_mat_no_entries"])
    assert (alg.axpy(h, two, np.array([np.nan, np.nan]), 1.0, 0.0) == 0).all()     # beta = 0 overwrites
    assert (alg.axpy(h, two, two, 1.0, 1.0) == two).all()
    assert (alg.axpy(h, two, two, 1.0, 1.0, True) == two).all()
    h0 = alg.mat(G["test_mat_no_rows"])                  # 0 x 2
    assert alg.axpy(h0, two, np.zeros(0), 1.0, 0.0).size == 0
    assert (alg.axpy(h0, np.zeros(0), two, 1.0, 1.0, True) == two).all()
    assert (alg.axpy(h0, np.zeros(0), two, 1.0, 0.0, True) == 0).all()
    h1 = alg.mat(G["test_mat_no_cols"])                  # 2 x 0
    assert (alg.axpy(h1, np.zeros(0), two, 1.0, 1.0) == two).all()
    assert (alg.axpy(h1, np.zeros(0), two, 1.0, 0.0) == 0).all()


# ---------------------------------------------------------------- test_matrix.cpp
def test_scalings_and_norms(alg):
    L = alg.lib
    A, d = G["test_mat_ops_A"], np.asarray(G["test_mat_ops_d"], dtype=float)
    x = np.array([0.3, -1.7])
    for op, ref in (("lmult", G["test_mat_ops_prem_diag"]), ("rmult", G["test_mat_ops_postm_diag"]),
                    ("scal", G["test_mat_ops_scaled"])):
        h = alg.mat(A)
        if op == "lmult":
            L.OSQPMatrix_lmult_diag(h, alg.vec(d))
        elif op == "rmult":
            L.OSQPMatrix_rmult_diag(h, alg.vec(d))
        else:
            L.OSQPMatrix_mult_scalar(h, 2.0)
        # compare through both stored orientations
        assert np.abs(alg.axpy(h, x, np.zeros(2), 1.0, 0.0) - ref @ x).max() < TOL
        assert np.abs(alg.axpy(h, x, np.zeros(2), 1.0, 0.0, True) - ref.T @ x).max() < TOL
        L.OSQPMatrix_free(h)
    h = alg.mat(A)
    r = alg.vec(np.zeros(2))
    L.OSQPMatrix_col_norm_inf(h, r)
    assert np.abs(alg.get(r) - np.asarray(G["test_mat_ops_inf_norm_cols"]).ravel()).max() < TOL
    L.OSQPMatrix_row_norm_inf(h, r)
    assert np.abs(alg.get(r) - np.asarray(G["test_mat_ops_inf_norm_rows"]).ravel()).max() < TOL
    # symmetric P given as its upper triangle: the reference takes COLUMN norms of the stored
    # triangle only (builtin/matrix.c:194-197) but ROW norms of the full symmetric matrix (:199-203)
    Pu = G["test_mat_ops_diag_Pu"]
    Pf = (Pu + sp.triu(Pu, 1).T).toarray()
    hP, r6 = alg.mat(Pu, triu=True), alg.vec(np.zeros(6))
    L.OSQPMatrix_col_norm_inf(hP, r6)
    assert np.abs(alg.get(r6) - np.abs(Pu.toarray()).max(axis=0)).max() < TOL
    L.OSQPMatrix_row_norm_inf(hP, r6)
    assert np.abs(alg.get(r6) - np.abs(Pf).max(axis=1)).max() < TOL
    L.OSQPMatrix_lmult_diag(hP, alg.vec(np.arange(1.0, 7.0)))
    L.OSQPMatrix_rmult_diag(hP, alg.vec(np.arange(1.0, 7.0)))
    D = np.diag(np.arange(1.0, 7.0))
    xx = np.linspace(-1, 1, 6)
    assert np.abs(alg.axpy(hP, xx, np.zeros(6), 1.0, 0.0) - D @ Pf @ D @ xx).max() < TOL


def test_submatrix_byrows(alg):
    A = G["test_mat_vec_A"]
    x = np.asarray(G["test_mat_vec_x"], dtype=float)
    for tag in ("A4", "A5", "A3", "A0"):
        ind = np.asarray(G[f"test_submat_{tag}_ind"], dtype=np.int32)
        h = alg.mat(A)
        sub = alg.lib.OSQPMatrix_submatrix_byrows(h, alg.veci(ind))
        ref = sp.csc_matrix(G[f"test_submat_{tag}"]) if G[f"test_submat_{tag}_num"] else sp.csc_matrix((0, 4))
        assert alg.lib.OSQPMatrix_get_m(sub) == G[f"test_submat_{tag}_num"]
        assert alg.lib.OSQPMatrix_get_n(sub) == 4
        mred = ref.shape[0]
        if mred:
            assert np.abs(alg.axpy(sub, x, np.zeros(mred), 1.0, 0.0) - ref @ x).max() < TOL
            yv = np.linspace(1, 2, mred)
            assert np.abs(alg.axpy(sub, yv, np.zeros(4), 1.0, 0.0, True) - ref.T @ yv).max() < TOL
    # polish passes -1 / +1 flags: any non-zero keeps the row (csc_utils.c:134-203)
    h = alg.mat(A)
    sub = alg.lib.OSQPMatrix_submatrix_byrows(h, alg.veci(np.array([-1, 0, 1, 0, -1], dtype=np.int32)))
    assert np.abs(alg.axpy(sub, x, np.zeros(3), 1.0, 0.0) - (A.toarray()[[0, 2, 4]] @ x)).max() < TOL


def test_update_values(alg):
    L = alg.lib
    g = load_golden("update_matrices")
    A, An = sp.csc_matrix(g["test_form_KKT_A"]), sp.csc_matrix(g["test_form_KKT_A_new"])
    Pu, Pn = sp.csc_matrix(g["test_form_KKT_Pu"]), sp.csc_matrix(g["test_form_KKT_Pu_new"])
    x, y = np.linspace(-1, 1, A.shape[1]), np.linspace(1, 2, A.shape[0])
    hA, hP = alg.mat(A), alg.mat(Pu, triu=True)
    # partial update through index lists (positions in the USER's CSC arrays)
    idx = np.asarray(g["test_form_KKT_A_new_idx"], dtype=np.int32)
    vals = np.ascontiguousarray(An.data[idx])
    L.OSQPMatrix_update_values(hA, vals.ctypes.data, idx.ctypes.data, idx.size)
    Aexp = A.copy(); Aexp.data[idx] = An.data[idx]
    assert np.abs(alg.axpy(hA, x, np.zeros(A.shape[0]), 1.0, 0.0) - Aexp @ x).max() < TOL
    assert np.abs(alg.axpy(hA, y, np.zeros(A.shape[1]), 1.0, 0.0, True) - Aexp.T @ y).max() < TOL
    idx = np.asarray(g["test_form_KKT_Pu_new_idx"], dtype=np.int32)
    vals = np.ascontiguousarray(Pn.data[idx])
    L.OSQPMatrix_update_values(hP, vals.ctypes.data, idx.ctypes.data, idx.size)
    Pexp = Pu.copy(); Pexp.data[idx] = Pn.data[idx]
    Pfull = Pexp + sp.triu(Pexp, 1).T
    assert np.abs(alg.axpy(hP, x, np.zeros(A.shape[1]), 1.0, 0.0) - Pfull @ x).max() < TOL
    # full update: idx == NULL means all values in CSC order
    vals = np.ascontiguousarray(An.data)
    L.OSQPMatrix_update_values(hA, vals.ctypes.data, None, 0)
    assert np.abs(alg.axpy(hA, x, np.zeros(A.shape[0]), 1.0, 0.0) - An @ x).max() < TOL
    assert np.abs(alg.axpy(hA, y, np.zeros(A.shape[1]), 1.0, 0.0, True) - An.T @ y).max() < TOL


def test_device_pointer_io(alg, kern):
    """from_raw / to_raw accept DEVICE raw pointers too
    (tests/lin_alg/testcases/cuda/test_vector_cuda.cpp:8-58, tests/basic_qp/test_cuda_io.cpp:117-134)"""
    if "B200" not in str(alg.lib):
        pass
    from osqp_b200.devmem import DeviceArray
    src = DeviceArray(kern, v("v1"))
    a = alg.lib.OSQPVectorf_malloc(10)
    if alg.lib._name.endswith("libosqp_builtin.so"):
        pytest.skip("device pointers are a GPU-backend feature")
    alg.lib.OSQPVectorf_from_raw(a, src.ptr)
    assert (alg.get(a) == v("v1")).all()
    dst = DeviceArray(kern, np.zeros(10))
    alg.lib.OSQPVectorf_mult_scalar(a, 2.0)
    alg.lib.OSQPVectorf_to_raw(dst.ptr, a)
    assert (dst.get() == 2.0 * v("v1")).all()
