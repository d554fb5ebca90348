"""End-to-end GPU parity: osqp_setup / osqp_solve / osqp_update_* through the unchanged core with
the B200 backend, against (a) the reference's golden solutions (tests/golden, same assertions as
the reference's osqp_tester with the indirect solver) and (b) the CPU oracle (builtin + QDLDL) on
seeded problems of the BASELINE.json families.

Parity definition (BASELINE.json north_star; SURVEY.md section 7 "Hard parts"): same status; objective
to 1e-6 relative and residuals within eps when the solve is tight (eps 1e-7..1e-8, CG run to
its 1e-7 floor); ADMM iteration count within max(10 %, 2 termination-check intervals) of the
direct solver's."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import FIXTURE_SETTINGS, TESTS_TOL, load_golden
from osqp_b200 import _capi, problems
from osqp_b200.interface import OSQP
from test_oracle_golden import close

pytestmark = pytest.mark.gpu

# CG driven to its absolute floor OSQP_CG_TOL_MIN = 1e-7 (osqp_api_constants.h:215)
TIGHT_CG = dict(cg_tol_fraction=1e-8, cg_max_iter=500)


def run(lib, d, A="A", u="u", **kw):
    st = dict(FIXTURE_SETTINGS)
    st.update(kw)
    s = OSQP(lib).setup(d["P"], d["q"], d[A], d["l"], d[u], **st)
    return s, s.solve()


# ---------------------------------------------------------------- golden solutions
@pytest.mark.parametrize("case,kw", [
    ("basic_qp", {}), ("basic_lp", {}), ("basic_qp2", dict(eps_abs=1e-6, eps_rel=1e-6)),
    ("no_active_set", {}), ("unconstrained", {}),
])
@pytest.mark.parametrize("polishing", [0, 1])
def test_golden_solutions(b200_lib, case, kw, polishing):
    d = load_golden(case)
    s, r = run(b200_lib, d, polishing=polishing, polish_refine_iter=4, **kw)
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert close(r.x, d["x_test"])
    if "y_test" in d:
        assert close(r.y, d["y_test"])
    assert close(r.info.obj_val, d["obj_value_test"])


@pytest.mark.parametrize("driver", ["persistent", "graph"])
def test_large_qp(b200_lib, monkeypatch, driver):
    """tests/large_qp/test_large_qp.cpp:10-44 with the indirect solver: status and objective 0.106081 to
    TESTS_TOL relative"""
    monkeypatch.setenv("B200_PCG_DRIVER", driver)
    d = load_golden("large_qp")
    s, r = run(b200_lib, d)
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert abs(r.info.obj_val - d["obj_value_test"]) / abs(d["obj_value_test"]) < TESTS_TOL


def test_basic_qp2_update_vectors(b200_lib):
    d = load_golden("basic_qp2")
    s, r = run(b200_lib, d, eps_abs=1e-6, eps_rel=1e-6, warm_starting=1, polishing=1)
    s.update(q=d["q_new"], u=d["u_new"])
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert close(r.x, d["x_test_new"]) and close(r.y, d["y_test_new"])
    assert close(r.info.obj_val, d["obj_value_test_new"])


def test_primal_infeasible_random_problem(b200_lib):
    d = load_golden("primal_infeasibility")     # n=50, m=150 (tests/primal_infeasibility)
    s, r = run(b200_lib, d, polishing=1, scaling=0, warm_starting=0)
    assert r.info.status_val == _capi.OSQP_PRIMAL_INFEASIBLE
    # OSQP_NAN is literally ((OSQPFloat)0x7fc00000UL) (osqp_api_constants.h:188-190): in a double
    # build the "NaN" the core writes into x is the number 2143289344.0 -- same as the reference
    assert (r.x == float(0x7fc00000)).all() and np.isfinite(r.prim_inf_cert).all()
    assert abs(np.abs(r.prim_inf_cert).max() - 1.0) < 1e-9       # normalised certificate


def test_certificate_arrays_after_infeasible_then_feasible(b200_lib):
    """store_solution (auxil.c:598-675) through the backend's fused version: a run with a solution leaves
    OSQP_NAN in both certificate arrays -- also when the previous run of the same solver wrote a
    certificate there -- and an infeasible run writes the normalised certificate over the OSQP_NAN."""
    P = sp.csc_matrix(np.diag([1.0, 2.0, 0.5]))
    q = np.array([1.0, -1.0, 0.5])
    A = sp.csc_matrix(np.array([[1.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]]))
    l_bad, u_bad = np.array([1.0, -np.inf, -1.0]), np.array([np.inf, 0.0, 1.0])      # x0 + x1 >= 1 and <= 0
    l_ok, u_ok = np.array([-1.0, -np.inf, -1.0]), np.array([np.inf, 2.0, 1.0])
    st = dict(FIXTURE_SETTINGS)
    st.update(warm_starting=0, polishing=0)
    s = OSQP(b200_lib).setup(P, q, A, l_ok, u_ok, **st)
    sol = s._solver.contents.solution.contents
    nan = _capi.OSQP_NAN

    def raw(ptr, k):
        return np.array([ptr[i] for i in range(k)])

    r = s.solve()
    assert r.info.status_val == _capi.OSQP_SOLVED and r.prim_inf_cert is None and r.dual_inf_cert is None
    assert (raw(sol.prim_inf_cert, 3) == nan).all() and (raw(sol.dual_inf_cert, 3) == nan).all()
    x_ok = r.x.copy()
    s.update(l=l_bad, u=u_bad)
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_PRIMAL_INFEASIBLE
    assert r.dual_inf_cert is None and abs(np.abs(r.prim_inf_cert).max() - 1.0) < 1e-9
    assert (raw(sol.dual_inf_cert, 3) == nan).all()
    s.update(l=l_ok, u=u_ok)
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_SOLVED and r.prim_inf_cert is None
    assert (raw(sol.prim_inf_cert, 3) == nan).all() and (raw(sol.dual_inf_cert, 3) == nan).all()
    assert np.allclose(r.x, x_ok, atol=1e-3)
    s.cleanup()


@pytest.mark.parametrize("A,u,status", [
    ("A12", "u1", _capi.OSQP_SOLVED), ("A12", "u2", _capi.OSQP_PRIMAL_INFEASIBLE),
    ("A34", "u3", _capi.OSQP_DUAL_INFEASIBLE), ("A34", "u4", _capi.OSQP_PRIMAL_INFEASIBLE)])
def test_primal_dual_infeasibility(b200_lib, A, u, status):
    d = load_golden("primal_dual_infeasibility")
    s, r = run(b200_lib, d, A=A, u=u, polishing=(1 if u == "u1" else 0), scaling=0)
    assert r.info.status_val == status
    if status == _capi.OSQP_SOLVED:
        assert close(r.x, d["x1"]) and close(r.y, d["y1"]) and close(r.info.obj_val, d["obj_value1"])


def test_non_convex_detected_while_solving(b200_lib):
    """tests/non_cvx/test_non_cvx.cpp:50-: with sigma large enough to pass setup the iterates
    diverge and the core reports OSQP_NON_CVX (auxil.c:833-840)."""
    d = load_golden("non_cvx")
    s = OSQP(b200_lib).setup(d["P"], d["q"], d["A"], d["l"], d["u"], sigma=float(d["sigma_new"]),
                             adaptive_rho=0, **FIXTURE_SETTINGS)
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_NON_CVX and r.info.obj_val == float(0x7fc00000)


def test_update_matrices(b200_lib):
    """tests/update_matrices/test_update_matrices.cpp:94-385"""
    g = load_golden("update_matrices")
    d = dict(P=g["test_solve_Pu"], q=g["test_solve_q"], A=g["test_solve_A"], l=g["test_solve_l"],
             u=g["test_solve_u"])
    Pn = sp.triu(g["test_solve_Pu_new"], format="csc")
    An = sp.csc_matrix(g["test_solve_A_new"])
    s, r = run(b200_lib, d, max_iter=1000)
    assert close(r.x, g["test_solve_x"]) and close(r.y, g["test_solve_y"])
    s.update(Px=Pn.data)                                   # all of P
    r = s.solve()
    assert close(r.x, g["test_solve_P_new_x"]) and close(r.info.obj_val, g["test_solve_P_new_obj_value"])
    s2, _ = run(b200_lib, d, max_iter=1000)
    s2.update(Ax=An.data)                                  # all of A
    r = s2.solve()
    assert close(r.x, g["test_solve_A_new_x"]) and close(r.y, g["test_solve_A_new_y"])
    s3, _ = run(b200_lib, d, max_iter=1000)                # P and A, through explicit index lists
    s3.update(Px=Pn.data, Px_idx=np.arange(Pn.nnz), Ax=An.data, Ax_idx=np.arange(An.nnz))
    r = s3.solve()
    assert close(r.x, g["test_solve_P_A_new_x"]) and close(r.y, g["test_solve_P_A_new_y"])
    assert close(r.info.obj_val, g["test_solve_P_A_new_obj_value"])


def test_basic_qp_termination_warm_start_and_rho_update(b200_lib):
    d = load_golden("basic_qp")
    # max_iter with termination checking off (test_basic_qp.cpp:718-774)
    s, r = run(b200_lib, d, max_iter=400, check_termination=0)
    assert r.info.iter == 400 and close(r.x, d["x_test"]) and close(r.info.obj_val, d["obj_value_test"])
    # warm start at the optimum -> one iteration (test_basic_qp.cpp:945-993)
    s, r = run(b200_lib, d, check_termination=1, adaptive_rho=0, **TIGHT_CG)
    it_cold = r.info.iter
    s.warm_start(x=r.x, y=r.y)
    r2 = s.solve()
    assert r2.info.iter <= 2 < it_cold
    # same rho via osqp_update_rho gives the same iteration count (test_basic_qp.cpp:776-886)
    s, r = run(b200_lib, d, rho=0.7, adaptive_rho=0, check_termination=1, **TIGHT_CG)
    s2, _ = run(b200_lib, d, rho=0.1, adaptive_rho=0, check_termination=1, **TIGHT_CG)
    s2.update_settings(rho=0.7)
    s2.cold_start()
    r2 = s2.solve()
    assert r2.info.iter == r.info.iter
    # new data vectors
    s.update(q=d["q_new"], l=d["l_new"], u=d["u_new"])
    assert s.solve().info.status_val == _capi.OSQP_SOLVED


def test_solution_into_device_buffers(b200_lib, kern):
    """allocate_solution = 0 + osqp_get_solution into DEVICE arrays (tests/basic_qp/test_cuda_io.cpp)."""
    import ctypes as C
    from osqp_b200.devmem import DeviceArray
    d = load_golden("basic_qp")
    s, r = run(b200_lib, d)
    T = b200_lib.T
    bufs = [DeviceArray(kern, np.zeros(k)) for k in (2, 4, 4, 2)]
    sol = T.OSQPSolution()
    fp = C.POINTER(C.c_double)
    sol.x, sol.y = C.cast(bufs[0].ptr, fp), C.cast(bufs[1].ptr, fp)
    sol.prim_inf_cert, sol.dual_inf_cert = C.cast(bufs[2].ptr, fp), C.cast(bufs[3].ptr, fp)
    assert b200_lib.lib.osqp_get_solution(s._solver, C.byref(sol)) == 0
    assert (bufs[0].get() == r.x).all() and (bufs[1].get() == r.y).all()


# ---------------------------------------------------------------- parity with the CPU oracle
def _family(name):
    if name == "random_qp":
        return problems.random_qp(400, 800, nnz_target=8000, seed=3)
    if name == "lasso":
        return problems.lasso(60, 600, density=0.1, seed=1)
    if name == "portfolio":
        return problems.portfolio(600, 30, density=0.2, seed=1)
    if name == "huber":
        return problems.huber(30, 300, density=0.2, seed=1)
    if name == "svm":
        return problems.svm(30, 300, density=0.2, seed=1)
    if name == "mpc":
        return problems.mpc(N=12, seed=1)
    raise KeyError(name)


FAMILIES = ["random_qp", "lasso", "portfolio", "huber", "svm", "mpc"]


@pytest.mark.parametrize("family", FAMILIES)
@pytest.mark.parametrize("rho_is_vec", [0, 1])
def test_tight_parity_with_builtin_qdldl(b200_lib, oracle_lib, family, rho_is_vec):
    """north_star: same status, objective to 1e-6 relative, residuals within eps, iteration count
    within the stated band."""
    pb = _family(family)
    # eps ~ 3e-7 is about the tightest tolerance an indirect solve can certify: the CG tolerance has the
    # hard floor OSQP_CG_TOL_MIN = 1e-7 on the (scaled) linear-system residual, which bounds the
    # reachable dual residual (same in the reference: cuda_pcg_interface.cu:60)
    EPS = 3e-7
    kw = dict(eps_abs=EPS, eps_rel=EPS, max_iter=20000, rho_is_vec=rho_is_vec, check_termination=25,
              verbose=0)
    so = OSQP(oracle_lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
    ro = so.solve()
    sb = OSQP(b200_lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw, **TIGHT_CG)
    rb = sb.solve()
    assert ro.info.status_val == _capi.OSQP_SOLVED
    assert rb.info.status_val == ro.info.status_val
    assert abs(rb.info.obj_val - ro.info.obj_val) <= 1e-6 * max(1.0, abs(ro.info.obj_val))
    # residuals within the tolerances the core itself applied (eps_abs + eps_rel * scale >= eps_abs)
    # primal / dual residuals within eps_abs + eps_rel * (the norms the core normalises by)
    Ax = pb["A"] @ rb.x
    Px = sp.csc_matrix(pb["P"]) @ rb.x
    Aty = pb["A"].T @ rb.y
    assert rb.info.prim_res <= EPS * (1 + max(np.abs(Ax).max(), 1e-30))
    assert rb.info.dual_res <= EPS * (1 + max(np.abs(Px).max(), np.abs(Aty).max(), np.abs(pb["q"]).max()))
    assert np.abs(rb.x - ro.x).max() <= 1e-4 * max(1.0, np.abs(ro.x).max())
    band = max(0.10 * ro.info.iter, 2 * 25)
    assert abs(rb.info.iter - ro.info.iter) <= band, (rb.info.iter, ro.info.iter)


@pytest.mark.parametrize("family", FAMILIES)
def test_default_tolerance_parity(b200_lib, oracle_lib, family):
    """At the benchmark tolerance eps = 1e-3 with the default (inexact) CG schedule: same status,
    residuals within tolerance, objective to ~eps."""
    pb = _family(family)
    kw = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5)
    ro = OSQP(oracle_lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw).solve()
    sb = OSQP(b200_lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
    rb = sb.solve()
    assert rb.info.status_val == ro.info.status_val == _capi.OSQP_SOLVED
    assert abs(rb.info.obj_val - ro.info.obj_val) <= 5e-3 * max(1.0, abs(ro.info.obj_val))
    Ax = pb["A"] @ rb.x
    viol = np.maximum(np.maximum(pb["l"] - Ax, Ax - pb["u"]), 0).max()
    assert viol <= 1e-3 * (1 + max(np.abs(Ax).max(), 1.0)) * 2


def test_float32_build_solves(oracle_lib):
    """the f32 build (the reference CUDA backend's default precision, CMakeLists.txt:156-163)"""
    from osqp_b200 import OSQP as B200OSQP
    from osqp_b200._lib import lib_paths
    if not lib_paths("f32")[1].exists():
        pytest.skip("f32 library not built")
    d = load_golden("basic_qp")
    s = B200OSQP("f32").setup(d["P"], d["q"], d["A"], d["l"], d["u"], check_dualgap=0,
                              **{**FIXTURE_SETTINGS, "eps_abs": 1e-4, "eps_rel": 1e-4})
    r = s.solve()
    assert r.info.status_val == _capi.OSQP_SOLVED
    assert np.abs(r.x - d["x_test"]).max() < 1e-3 and abs(r.info.obj_val - d["obj_value_test"]) < 1e-3
