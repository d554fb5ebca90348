"""Pin the CPU oracle (oracle/_ref/libosqp_builtin.so = unmodified reference core + builtin
backend + LOCAL QDLDL restatement) against every golden answer the reference's own osqp_tester
holds for this path, and the QDLDL restatement against scipy splu."""
import ctypes as C

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from conftest import FIXTURE_SETTINGS, ROOT, TESTS_TOL, load_golden
from osqp_b200 import _capi
from osqp_b200.interface import OSQP


def close(a, b, tol=TESTS_TOL):
    """the reference's comparison: inf-norm of the difference, relative for large solutions
    (tests/basic_qp2/test_basic_qp2.cpp:50-63)"""
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    if a.size == 0:
        return True
    return np.abs(a - b).max() < tol * max(1.0, np.abs(b).max())


def solve(lib, d, A="A", u="u", **kw):
    st = dict(FIXTURE_SETTINGS)
    st.update(kw)
    s = OSQP(lib).setup(d["P"], d["q"], d[A], d["l"], d[u], **st)
    return s, s.solve()


@pytest.mark.parametrize("case,kw", [
    ("basic_qp", {}), ("basic_lp", {}), ("basic_qp2", dict(eps_abs=1e-6, eps_rel=1e-6)),
    ("no_active_set", {}), ("unconstrained", {}),
])
@pytest.mark.parametrize("polishing", [0, 1])
def test_reference_golden_solutions(oracle_lib, case, kw, polishing):
    d = load_golden(case)
    s, r = solve(oracle_lib, d, polishing=polishing, polish_refine_iter=4, **kw)
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert close(r.x, d["x_test"])
    if "y_test" in d:
        assert close(r.y, d["y_test"])
    assert close(r.info.obj_val, d["obj_value_test"]) * max(1.0, abs(d["obj_value_test"]))


def test_large_qp(oracle_lib):
    """tests/large_qp/test_large_qp.cpp:10-44 (n = 160, m = 270): status and objective 0.106081 to TESTS_TOL
    relative"""
    d = load_golden("large_qp")
    s, r = solve(oracle_lib, d)
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert abs(r.info.obj_val - d["obj_value_test"]) / abs(d["obj_value_test"]) < TESTS_TOL


def test_basic_qp2_update(oracle_lib):
    d = load_golden("basic_qp2")
    s, r = solve(oracle_lib, d, eps_abs=1e-6, eps_rel=1e-6, warm_starting=1, polishing=1)
    s.update(q=d["q_new"], u=d["u_new"])
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert close(r.x, d["x_test_new"])
    assert close(r.y, d["y_test_new"])
    assert close(r.info.obj_val, d["obj_value_test_new"]) * abs(d["obj_value_test_new"])


def test_primal_infeasibility(oracle_lib):
    d = load_golden("primal_infeasibility")
    s, r = solve(oracle_lib, d, polishing=1, scaling=0, warm_starting=0)
    assert r.info.status_val == _capi.OSQP_PRIMAL_INFEASIBLE
    # the wrapper hands out a certificate only where the core wrote one (OSQP_NAN marker elsewhere)
    assert r.dual_inf_cert is None and abs(np.abs(r.prim_inf_cert).max() - 1.0) < 1e-9
    assert (r.x == _capi.OSQP_NAN).all()


@pytest.mark.parametrize("A,u,status", [
    ("A12", "u1", _capi.OSQP_SOLVED), ("A12", "u2", _capi.OSQP_PRIMAL_INFEASIBLE),
    ("A34", "u3", _capi.OSQP_DUAL_INFEASIBLE), ("A34", "u4", _capi.OSQP_PRIMAL_INFEASIBLE)])
def test_primal_dual_infeasibility(oracle_lib, A, u, status):
    d = load_golden("primal_dual_infeasibility")
    s, r = solve(oracle_lib, d, A=A, u=u, polishing=(1 if u == "u1" else 0), scaling=0)
    assert r.info.status_val == status
    if status == _capi.OSQP_SOLVED:
        assert close(r.x, d["x1"])
        assert close(r.y, d["y1"])
        assert close(r.info.obj_val, d["obj_value1"])


def test_update_matrices(oracle_lib):
    g = load_golden("update_matrices")
    d = dict(P=g["test_solve_Pu"], q=g["test_solve_q"], A=g["test_solve_A"], l=g["test_solve_l"],
             u=g["test_solve_u"])
    s, r = solve(oracle_lib, d, max_iter=1000)
    assert close(r.x, g["test_solve_x"])
    assert close(r.y, g["test_solve_y"])
    Pn = sp.triu(g["test_solve_Pu_new"], format="csc")
    s.update(Px=Pn.data)
    r = s.solve()
    assert close(r.x, g["test_solve_P_new_x"])
    assert close(r.info.obj_val, g["test_solve_P_new_obj_value"])
    s.update(Ax=sp.csc_matrix(g["test_solve_A_new"]).data)
    r = s.solve()
    assert close(r.x, g["test_solve_P_A_new_x"])
    assert close(r.y, g["test_solve_P_A_new_y"])


def test_non_convex_detected_at_setup(oracle_lib):
    """tests/non_cvx/test_non_cvx.cpp:9-48: QDLDL's positive-pivot count flags P + sigma I
    indefinite -> OSQP_NONCVX_ERROR (4); pins the sign convention of the QDLDL restatement."""
    from osqp_b200.interface import OSQPError
    d = load_golden("non_cvx")
    with pytest.raises(OSQPError) as e:
        OSQP(oracle_lib).setup(d["P"], d["q"], d["A"], d["l"], d["u"], sigma=1e-6, adaptive_rho=0,
                               **{k: v for k, v in FIXTURE_SETTINGS.items()})
    assert e.value.code == 4


# ------------------------------------------------------------------ QDLDL restatement vs splu
def _qdldl():
    lib = C.CDLL(str(ROOT / "oracle" / "_ref" / "libqdldl_oracle.so"))
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    lib.QDLDL_etree.argtypes = [C.c_int, ip, ip, ip, ip, ip]
    lib.QDLDL_etree.restype = C.c_int
    lib.QDLDL_factor.argtypes = [C.c_int, ip, ip, fp, ip, ip, fp, fp, fp, ip, ip, ip, ip, fp]
    lib.QDLDL_factor.restype = C.c_int
    lib.QDLDL_solve.argtypes = [C.c_int, ip, ip, fp, fp, fp]
    return lib


def _ldl_solve(lib, K, b):
    Ku = sp.triu(K, format="csc")
    Ku.sort_indices()
    n = Ku.shape[0]
    Ap, Ai = Ku.indptr.astype(np.int32), Ku.indices.astype(np.int32)
    Ax = Ku.data.astype(np.float64)
    ip, fp = C.POINTER(C.c_int), C.POINTER(C.c_double)
    P = lambda a: a.ctypes.data_as(ip if a.dtype == np.int32 else fp)
    work, Lnz, etree = (np.zeros(n, np.int32) for _ in range(3))
    tot = lib.QDLDL_etree(n, P(Ap), P(Ai), P(work), P(Lnz), P(etree))
    assert tot >= 0
    Lp, Li, Lx = np.zeros(n + 1, np.int32), np.zeros(max(tot, 1), np.int32), np.zeros(max(tot, 1))
    D, Dinv, fwork = np.zeros(n), np.zeros(n), np.zeros(n)
    bwork, iwork = np.zeros(n, np.int32), np.zeros(3 * n, np.int32)
    npos = lib.QDLDL_factor(n, P(Ap), P(Ai), P(Ax), P(Lp), P(Li), P(Lx), P(D), P(Dinv), P(Lnz),
                            P(etree), P(bwork), P(iwork), P(fwork))
    x = np.array(b, dtype=np.float64)
    lib.QDLDL_solve(n, P(Lp), P(Li), P(Lx), P(Dinv), P(x))
    return x, npos, D


def test_qdldl_restatement_matches_reference_kkt_solution(oracle_lib):
    """tests/solve_linsys/generate_problem.py: KKT x = rhs with the answer from scipy."""
    g = load_golden("solve_linsys")
    x, npos, D = _ldl_solve(_qdldl(), g["test_solve_KKT_KKT"], g["test_solve_KKT_rhs"])
    assert npos == g["test_solve_KKT_n"]                 # n positive pivots, m negative
    n, rho = g["test_solve_KKT_n"], g["test_solve_KKT_rho"]
    # the golden vector holds (x~, z~ = b2 + nu / rho), the LinSysSolver.solve contract
    # (algebra/_common/lin_sys/qdldl/qdldl_interface.c:441-456)
    x[n:] = g["test_solve_KKT_rhs"][n:] + x[n:] / rho
    assert np.abs(x - g["test_solve_KKT_x"]).max() < 1e-9


@pytest.mark.parametrize("n,m,seed", [(30, 50, 0), (200, 300, 1), (1, 0, 2), (40, 0, 3)])
def test_qdldl_restatement_vs_splu(oracle_lib, n, m, seed):
    rng = np.random.default_rng(seed)
    M = sp.random(n, n, density=min(1.0, 5.0 / n), random_state=seed, format="csc")
    Pm = (M @ M.T + 1e-3 * sp.eye(n)).tocsc()
    A = sp.random(m, n, density=min(1.0, 4.0 / max(n, 1)), random_state=seed + 1, format="csc")
    K = sp.bmat([[Pm, A.T], [A, -10.0 * sp.eye(m)]], format="csc") if m else Pm
    b = rng.standard_normal(n + m)
    x, npos, D = _ldl_solve(_qdldl(), K, b)
    assert npos == n and (D[n:] < 0).all() if m else npos == n
    xr = sla.splu(sp.csc_matrix(K)).solve(b)
    assert np.abs(x - xr).max() < 1e-8 * max(1.0, np.abs(xr).max())


def test_qdldl_rejects_non_triu_and_zero_pivot(oracle_lib):
    lib = _qdldl()
    ip = C.POINTER(C.c_int)
    P = lambda a: a.ctypes.data_as(ip)
    # entry below the diagonal -> -1
    Ap, Ai = np.array([0, 2, 3], np.int32), np.array([0, 1, 1], np.int32)
    w, l, e = (np.zeros(2, np.int32) for _ in range(3))
    assert lib.QDLDL_etree(2, P(Ap), P(Ai), P(w), P(l), P(e)) == -1
    # empty column -> -1
    Ap, Ai = np.array([0, 1, 1], np.int32), np.array([0], np.int32)
    assert lib.QDLDL_etree(2, P(Ap), P(Ai), P(w), P(l), P(e)) == -1
