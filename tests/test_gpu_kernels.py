"""GPU tests at the C-ABI (include/osqp_b200.h): fused ADMM kernels against a numpy restatement of
src/auxil.c, the PCG solver (both device-resident drivers) against the reference's solve_linsys
golden vector and scipy, and size-independent properties at the full BASELINE size."""
import ctypes as C
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as sla

from conftest import load_golden
from osqp_b200 import problems
from osqp_b200.devmem import DeviceArray, csr_to_device

pytestmark = pytest.mark.gpu
F = 8


def dev(k, a):
    return DeviceArray(k, np.ascontiguousarray(a, dtype=np.float64))


# ------------------------------------------------------------------ fused ADMM step kernels
@pytest.mark.parametrize("rho_is_vec", [0, 1])
@pytest.mark.parametrize("n,m", [(7, 11), (1000, 3), (0, 5), (2049, 4097)])
def test_fused_rhs_and_update_match_auxil_c(kern, rho_is_vec, n, m):
    """compute_rhs (auxil.c:136-158) and update_x/z/y (auxil.c:172-229), term by term."""
    k = kern
    rng = np.random.default_rng(n + m)
    x_prev, q, xt = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(n)
    z_prev, y, zt = rng.standard_normal(m), rng.standard_normal(m), rng.standard_normal(m)
    l = rng.standard_normal(m) - 1.0
    u = l + rng.random(m) * 2
    l[::3] = -1e30
    u[1::4] = 1e30
    rho_vec = np.where(rng.random(m) < 0.3, 100.0, 0.1)
    rho_inv_vec = 1.0 / rho_vec
    rho, sigma, alpha = 0.1, 1e-6, 1.6
    rho_inv = 1.0 / rho
    # ---- compute_rhs
    d_xz = dev(k, np.zeros(n + m))
    dq, dxp, dzp, dy = dev(k, q), dev(k, x_prev), dev(k, z_prev), dev(k, y)
    drv, driv = dev(k, rho_vec), dev(k, rho_inv_vec)
    k.b200_admm_compute_rhs(d_xz.ptr, d_xz.offset(n), dxp.ptr, dq.ptr, dzp.ptr, dy.ptr,
                            driv.ptr if rho_is_vec else None, rho_inv, sigma, n, m)
    out = d_xz.get()
    ref_x = sigma * x_prev + (-1.0) * q
    ref_z = (-1.0) * (rho_inv_vec * y) + 1.0 * z_prev if rho_is_vec else 1.0 * z_prev + (-rho_inv) * y
    assert np.allclose(out[:n], ref_x, rtol=1e-15, atol=1e-15)
    assert np.allclose(out[n:], ref_z, rtol=1e-14, atol=1e-15)
    # ---- update_x / update_z / update_y
    dx, ddx, dz, ddy = dev(k, np.zeros(n)), dev(k, np.zeros(n)), dev(k, np.zeros(m)), dev(k, np.zeros(m))
    dxt, dzt, dl, du = dev(k, xt), dev(k, zt), dev(k, l), dev(k, u)
    k.b200_admm_update_xzy(dx.ptr, ddx.ptr, dz.ptr, dy.ptr, ddy.ptr, dxt.ptr, dzt.ptr, dxp.ptr, dzp.ptr,
                           dl.ptr, du.ptr, drv.ptr if rho_is_vec else None, driv.ptr if rho_is_vec else None,
                           rho, rho_inv, alpha, n, m)
    x_new = alpha * xt + (1 - alpha) * x_prev
    rinv = rho_inv_vec if rho_is_vec else rho_inv
    rr = rho_vec if rho_is_vec else rho
    z_new = np.minimum(np.maximum(alpha * zt + (1 - alpha) * z_prev + rinv * y, l), u)
    dy_ref = rr * (alpha * zt + (1 - alpha) * z_prev - z_new)
    assert np.allclose(dx.get(), x_new, rtol=1e-14, atol=1e-15)
    assert np.allclose(ddx.get(), x_new - x_prev, rtol=1e-13, atol=1e-14)
    assert np.allclose(dz.get(), z_new, rtol=1e-13, atol=1e-14)
    assert np.allclose(ddy.get(), dy_ref, rtol=1e-12, atol=1e-12)
    assert np.allclose(dy.get(), y + dy_ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("scaled", [0, 1])
@pytest.mark.parametrize("do_primal,do_dual", [(1, 1), (1, 0), (0, 1)])
@pytest.mark.parametrize("n,m", [(13, 29), (300, 70000), (4000, 7)])
def test_fused_infeasibility_pretest_scalars(kern, scaled, do_primal, do_dual, n, m):
    """the five scalars that decide whether is_primal_infeasible / is_dual_infeasible (auxil.c:460-585) can
    fire, from ONE kernel; delta_y is projected on the polar of the recession cone in place only when the
    primal test is asked for"""
    k = kern
    rng = np.random.default_rng(11 * n + m)
    dy, dx, q = rng.standard_normal(m), rng.standard_normal(n), rng.standard_normal(n)
    l = rng.standard_normal(m) - 1.0
    u = l + rng.random(m) * 2
    l[::3] = -1e30
    u[1::4] = 1e30
    E, D = rng.random(m) + 0.5, rng.random(n) + 0.5
    infval = 1e30 * 1e-4
    d_dy, d_l, d_u, d_dx, d_q, d_E, d_D = (dev(k, v) for v in (dy, l, u, dx, q, E, D))
    out = np.full(5, -1.0)
    k.b200_admm_infeas_scalars(d_dy.ptr, d_l.ptr, d_u.ptr, d_E.ptr if scaled else None, d_dx.ptr,
                               d_D.ptr if scaled else None, d_q.ptr, infval, n, m, do_primal, do_dual, out.ctypes.data)
    yp = dy.copy()
    both = (u > infval) & (l < -infval)
    yp[both] = 0
    up = (u > infval) & ~both
    yp[up] = np.minimum(yp[up], 0)
    lo = (l < -infval) & ~(u > infval)
    yp[lo] = np.maximum(yp[lo], 0)
    Ev, Dv = (E if scaled else np.ones(m)), (D if scaled else np.ones(n))
    ref = [np.abs(Ev * yp).max(), u @ np.maximum(yp, 0), l @ np.minimum(yp, 0), np.abs(Dv * dx).max(), q @ dx]
    if do_primal:
        assert np.array_equal(d_dy.get(), yp)                     # projected in place, bit for bit
        assert abs(out[0] - ref[0]) <= 1e-12 * ref[0]
        for s in (1, 2):
            assert abs(out[s] - ref[s]) <= 1e-9 * max(1.0, abs(ref[s]))
    else:
        assert np.array_equal(d_dy.get(), dy) and (out[:3] == 0).all()
    if do_dual:
        assert abs(out[3] - ref[3]) <= 1e-12 * ref[3] and abs(out[4] - ref[4]) <= 1e-9 * max(1.0, abs(ref[4]))
    else:
        assert (out[3:] == 0).all()


@pytest.mark.parametrize("scaled", [0, 1])
@pytest.mark.parametrize("n,m", [(13, 29), (5000, 1), (300, 70000)])
def test_fused_residual_reductions(kern, scaled, n, m):
    """every scalar of compute_prim_res / compute_dual_res / compute_obj_val_dual_gap /
    compute_*_tol (auxil.c:231-458) from ONE kernel"""
    k = kern
    rng = np.random.default_rng(7 * n + m)
    x, q, Px, Aty = (rng.standard_normal(n) for _ in range(4))
    y, z, Ax = (rng.standard_normal(m) for _ in range(3))
    y[::5] *= 1e-17                                # inside the dead zone
    l = rng.standard_normal(m) - 1.0
    u = l + rng.random(m) * 2
    l[::3] = -1e30
    u[1::4] = 1e30
    Einv, Dinv = rng.random(m) + 0.5, rng.random(n) + 0.5
    infval, dead = 1e30 * 1e-4, 1e-15
    out = np.zeros(17)
    args = [dev(k, v) for v in (x, y, z, Ax, Px, Aty, q, l, u)]
    dE, dD = dev(k, Einv), dev(k, Dinv)
    k.b200_admm_residuals(*[a.ptr for a in args], dE.ptr if scaled else None, dD.ptr if scaled else None,
                          infval, dead, n, m, out.ctypes.data)
    E = Einv if scaled else np.ones(m)
    D = Dinv if scaled else np.ones(n)
    yp = y.copy()
    both = (u > infval) & (l < -infval)
    yp[both] = 0
    up = (u > infval) & ~both
    yp[up] = np.minimum(yp[up], 0)
    lo = (l < -infval) & ~(u > infval)
    yp[lo] = np.maximum(yp[lo], 0)
    yp[np.abs(yp) < dead] = 0
    ref = [np.abs(Ax - z).max(), np.abs(E * (Ax - z)).max(), np.abs(z).max(), np.abs(E * z).max(),
           np.abs(Ax).max(), np.abs(E * Ax).max(), u @ np.maximum(yp, 0) + l @ np.minimum(yp, 0),
           np.abs(q + Px + Aty).max(), np.abs(D * (q + Px + Aty)).max(), np.abs(q).max(),
           np.abs(D * q).max(), np.abs(Px).max(), np.abs(D * Px).max(), np.abs(Aty).max(),
           np.abs(D * Aty).max(), Px @ x, q @ x]
    scale = np.maximum(1.0, np.abs(ref))
    assert (np.abs(out - ref) <= 1e-10 * scale * max(1.0, 1e-26 * np.abs(ref[6]) * 0 + 1)).all() or \
        np.allclose(out, ref, rtol=1e-9, atol=1e-9 * np.abs(ref[6]) if abs(ref[6]) > 1e20 else 1e-9)


# ------------------------------------------------------------------ PCG at the vtable level
def _full_P(Pu, n):
    Pu = sp.triu(sp.csc_matrix(Pu), format="csr")
    return (Pu + sp.triu(Pu, 1).T + sp.eye(n, format="csr") * 1e-300).tocsr()


def _pcg(k, P, A, sigma, rho, rho_vec=None, polishing=0, driver=None):
    if driver:
        os.environ["B200_PCG_DRIVER"] = driver
    n, m = P.shape[0], A.shape[0]
    hP, hA, hAt = csr_to_device(k, _full_P(P, n)), csr_to_device(k, sp.csr_matrix(A)), csr_to_device(k, sp.csr_matrix(A.T))
    s = k.b200_pcg_create(hP, hA, hAt, n, m)
    os.environ.pop("B200_PCG_DRIVER", None)
    drv = dev(k, rho_vec) if rho_vec is not None else None
    k.b200_pcg_configure(s, sigma, rho, drv.ptr if drv else None, 1, polishing)
    k.b200_pcg_refresh_matrices(s)
    k.b200_pcg_refresh_precond(s)
    return s, (hP, hA, hAt, drv)


@pytest.mark.parametrize("driver", ["persistent", "graph"])
def test_pcg_reference_known_answer(kern, driver):
    """The reference's (disabled) vtable-level test tests/solve_linsys/test_solve_linsys.h:8-59:
    prim_res = dual_res = 1e-7 forces a tight CG; solve(rhs, admm_iter = 2) must return
    (x~, z~ = b2 + nu / rho) of the KKT system."""
    k = kern
    g = load_golden("solve_linsys")
    n, m = int(g["test_solve_KKT_n"]), int(g["test_solve_KKT_m"])
    s, keep = _pcg(k, g["test_solve_KKT_Pu"], g["test_solve_KKT_A"], float(g["test_solve_KKT_sigma"]),
                   float(g["test_solve_KKT_rho"]), driver=driver)
    b = dev(k, g["test_solve_KKT_rhs"])
    for _ in range(3):                     # warm-started repeats converge to the same answer
        k.b200_copy_in(b.ptr, np.ascontiguousarray(g["test_solve_KKT_rhs"]).ctypes.data, (n + m) * F)
        assert k.b200_pcg_solve(s, b.ptr, 2, 1e-7, 1e-7, 500, 0.15, 10) == 0
    assert np.abs(b.get() - g["test_solve_KKT_x"]).max() < 1e-5
    k.b200_pcg_destroy(s)


@pytest.mark.parametrize("driver", ["persistent", "graph"])
@pytest.mark.parametrize("case", ["scalar_rho", "rho_vec", "polish", "unconstrained", "long_rows", "mid_rows"])
def test_pcg_against_direct_solve(kern, driver, case):
    k = kern
    rng = np.random.default_rng(3)
    n, m = 300, 500
    if case == "long_rows":
        n, m = 6000, 40
        A = sp.random(m, n, density=0.9, format="csr", random_state=4)        # rows of ~5400 entries
    elif case == "mid_rows":
        # rows of A between half a tile and a tile (1024 < len <= 2048): single CTA-wide chunks; the columns
        # of A (rows of A') stay short, so both kinds of tile occur in one solve
        n, m = 2500, 60
        A = sp.random(m, n, density=0.6, format="csr", random_state=4)        # rows of ~1500 entries
    elif case == "unconstrained":
        m = 0
        A = sp.csr_matrix((0, n))
    else:
        A = sp.random(m, n, density=0.03, format="csr", random_state=4)
    M = sp.random(n, n, density=0.01, format="csr", random_state=5)
    P = (M @ M.T + 0.1 * sp.eye(n)).tocsc()
    sigma, rho = 1e-3, 0.7
    rho_vec = np.where(rng.random(m) < 0.3, 50.0, 0.7) if case == "rho_vec" else None
    polishing = 1 if case == "polish" else 0
    if polishing:
        sigma, rho = 1e-3, 1.0 / 1e-3                                        # delta, 1 / delta
    R = sp.diags(rho_vec) if rho_vec is not None else rho * sp.eye(m)
    K = (P + sigma * sp.eye(n) + A.T @ R @ A).tocsc()
    s, keep = _pcg(k, P, A, sigma, rho, rho_vec, polishing, driver)
    rhs = rng.standard_normal(n + m)
    b = dev(k, rhs)
    for it in range(4):
        k.b200_copy_in(b.ptr, rhs.ctypes.data, (n + m) * F)
        assert k.b200_pcg_solve(s, b.ptr, 2, 1e-9, 1e-9, 3000, 0.15, 10) == 0
    out = b.get()
    x_ref = sla.spsolve(K, rhs[:n] + A.T @ (R @ rhs[n:]))
    # polishing solves only to 1e-5 ||rhs|| (OSQP_CG_POLISH_TOL, cuda_pcg_interface.cu:40)
    xtol = 1e-3 if polishing else 1e-5
    assert np.abs(out[:n] - x_ref).max() < xtol * max(1.0, np.abs(x_ref).max())
    if m:
        # second half of the contract, evaluated at the returned x: z~ = A x~, or (A x - b2) / delta
        z_exp = (A @ out[:n] - rhs[n:]) * rho if polishing else A @ out[:n]
        assert np.abs(out[n:] - z_exp).max() < 1e-9 * max(1.0, np.abs(z_exp).max())
    tot, ns, last = C.c_longlong(0), C.c_longlong(0), C.c_int(0)
    k.b200_pcg_stats(s, C.byref(tot), C.byref(ns), C.byref(last), None, None)
    assert ns.value == 4 and tot.value > 0 and last.value <= 5      # warm start: last solve is (nearly) free
    k.b200_pcg_destroy(s)


def test_pcg_tolerance_schedule_and_determinism(kern):
    """compute_tolerance (cuda_pcg_interface.cu:32-64): first ADMM iteration uses
    cg_tol_fraction * ||rhs||_inf; later ones max(min(lambda sqrt(pr dr), eps_prev), 1e-7)."""
    k = kern
    rng = np.random.default_rng(5)
    n, m = 400, 700
    A = sp.random(m, n, density=0.02, format="csr", random_state=1)
    P = sp.diags(rng.random(n) + 0.1).tocsc()
    rhs = rng.standard_normal(n + m)
    outs = []
    for rep in range(2):
        s, keep = _pcg(k, P, A, 1e-6, 0.1)
        b = dev(k, rhs)
        eps, its = C.c_double(0), C.c_int(0)
        k.b200_pcg_solve(s, b.ptr, 1, 0.0, 0.0, 20, 0.15, 10)
        k.b200_pcg_stats(s, None, None, C.byref(its), C.byref(eps), None)
        rhs_red = rhs[:n] + A.T @ (0.1 * rhs[n:])
        assert abs(eps.value - 0.15 * np.abs(rhs_red).max()) < 1e-12 * max(1.0, eps.value)
        first = b.get()
        k.b200_copy_in(b.ptr, rhs.ctypes.data, (n + m) * F)
        k.b200_pcg_solve(s, b.ptr, 2, 1e-2, 4e-2, 20, 0.15, 10)
        k.b200_pcg_stats(s, None, None, C.byref(its), C.byref(eps), None)
        assert abs(eps.value - min(0.15 * np.sqrt(1e-2 * 4e-2), 0.15 * np.abs(rhs_red).max())) < 1e-15
        outs.append((first, b.get()))
        k.b200_pcg_destroy(s)
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()   # bit-reproducible


# ------------------------------------------------------------------ full BASELINE size properties
def test_full_size_spmv_properties(kern):
    """BASELINE configs[1] size (n = m = 1.2e6, 1.14e7 nnz): adjoint identity <A x, y> = <x, A'y>,
    linearity, and agreement with scipy on a row sample."""
    k = kern
    pb = problems.lasso(100_000, 1_000_000, density=1e-4, seed=1)
    A = sp.csr_matrix(pb["A"])
    At = sp.csr_matrix(pb["A"].T)
    m, n = A.shape
    hA, hAt = csr_to_device(k, A), csr_to_device(k, At)
    rng = np.random.default_rng(0)
    x, x2, y = rng.standard_normal(n), rng.standard_normal(n), rng.standard_normal(m)
    dx, dx2, dyv = dev(k, x), dev(k, x2), dev(k, y)
    dAx, dAx2, dAty, dsum = dev(k, np.zeros(m)), dev(k, np.zeros(m)), dev(k, np.zeros(n)), dev(k, np.zeros(n))
    k.b200_csr_spmv(hA, dx.ptr, dAx.ptr, 1.0, 0.0)
    k.b200_csr_spmv(hAt, dyv.ptr, dAty.ptr, 1.0, 0.0)
    lhs = k.b200_vec_dot(dAx.ptr, dyv.ptr, m)
    rhs = k.b200_vec_dot(dx.ptr, dAty.ptr, n)
    assert abs(lhs - rhs) <= 1e-10 * max(abs(lhs), 1.0)
    # linearity: A (2 x + 3 x2) == 2 A x + 3 A x2
    k.b200_vec_add_scaled(dsum.ptr, 2.0, dx.ptr, 3.0, dx2.ptr, n)
    k.b200_csr_spmv(hA, dsum.ptr, dAx2.ptr, 1.0, 0.0)
    k.b200_csr_spmv(hA, dx2.ptr, dAx.ptr, 3.0, 2.0)           # y = 3 A x2 + 2 (A x)
    assert k.b200_vec_norm_inf_diff(dAx.ptr, dAx2.ptr, m) < 1e-10
    rows = rng.integers(0, m, 2000)
    assert np.abs(dAx2.get()[rows] - (A[rows] @ (2 * x + 3 * x2))).max() < 1e-10
    k.b200_csr_destroy(hA)
    k.b200_csr_destroy(hAt)


@pytest.mark.parametrize("shape,density", [((300, 70), 0.1), ((2000, 500), 0.02), ((64, 4000), 0.3), ((50, 50), 0.0)])
def test_device_transpose_matches_host_order(kern, shape, density):
    """b200_csr_transpose: CSR of the transpose built on the device (count / scan / scatter / per-row
    rank sort) must be bit-identical to the sorted host transpose, whatever order the atomics were
    served in, and the index map must send every source entry to its copy."""
    k = kern
    rng = np.random.default_rng(7)
    M = sp.random(shape[0], shape[1], density=density, format="csr", random_state=3,
                  data_rvs=lambda s: rng.standard_normal(s))
    if density == 0.0:
        M = sp.csr_matrix(([1.5, -2.5], ([3, 3], [7, 1])), shape=shape)    # nearly empty: empty rows on both sides
    M.sort_indices()
    h = csr_to_device(k, M)
    d_map = C.c_void_p()
    ht = k.b200_csr_transpose(h, C.byref(d_map))
    assert ht, "device transpose refused a matrix with short rows"
    Mt = sp.csr_matrix(M.T)
    Mt.sort_indices()
    rp = np.zeros(Mt.shape[0] + 1, dtype=np.int32)
    ci = np.zeros(max(Mt.nnz, 1), dtype=np.int32)
    vx = np.zeros(max(Mt.nnz, 1), dtype=np.float64)
    assert k.b200_csr_download(ht, rp.ctypes.data, ci.ctypes.data, vx.ctypes.data) == 0
    assert np.array_equal(rp, Mt.indptr) and np.array_equal(ci[:Mt.nnz], Mt.indices)
    assert np.array_equal(vx[:Mt.nnz], Mt.data)                       # values moved, not recomputed
    mp = np.zeros(M.nnz, dtype=np.int32)
    assert k.b200_copy_out(mp.ctypes.data, d_map, M.nnz * 4) == 0
    assert np.array_equal(vx[mp], M.data) and np.unique(mp).size == M.nnz
    # twice the same bits
    ht2 = k.b200_csr_transpose(h, None)
    ci2, vx2 = np.zeros_like(ci), np.zeros_like(vx)
    assert k.b200_csr_download(ht2, rp.ctypes.data, ci2.ctypes.data, vx2.ctypes.data) == 0
    assert np.array_equal(ci, ci2) and np.array_equal(vx, vx2)
    k.b200_free(d_map)
    for hh in (h, ht, ht2):
        k.b200_csr_destroy(hh)


def test_device_transpose_declines_over_long_rows(kern):
    """A destination row longer than the rank-sort limit keeps the host path (NULL, no crash)."""
    k = kern
    M = sp.csr_matrix(np.ones((5000, 3)))           # transpose has rows of 5000 entries
    h = csr_to_device(k, M)
    assert not k.b200_csr_transpose(h, None)
    k.b200_csr_destroy(h)


@pytest.mark.parametrize("n,density,drop_diag", [(200, 0.05, False), (300, 0.02, True), (64, 0.5, True), (500, 0.0, True)])
def test_device_symmetric_expansion(kern, n, density, drop_diag):
    """b200_csr_symmetric_from_triu: full symmetric CSR with a structurally full diagonal, expanded on
    the device from the upper-triangular CSC; both index maps must address the right copies."""
    k = kern
    rng = np.random.default_rng(5)
    M = sp.random(n, n, density=density, format="csc", random_state=9, data_rvs=lambda s: rng.standard_normal(s))
    U = sp.triu(M + M.T + sp.diags(rng.standard_normal(n)), format="csc")
    if drop_diag:       # structurally missing diagonal entries get an explicit zero
        U = U.tolil()
        for i in range(0, n, 3):
            U[i, i] = 0.0
        U = U.tocsc()
        U.eliminate_zeros()
    if density == 0.0:
        U = sp.csc_matrix(([2.0, -1.0, 4.0], ([0, 1, 5], [0, 7, 5])), shape=(n, n))
    U.sort_indices()
    p = np.ascontiguousarray(U.indptr, dtype=np.int32)
    i = np.ascontiguousarray(U.indices, dtype=np.int32)
    x = np.ascontiguousarray(U.data, dtype=np.float64)
    mu, ml = C.c_void_p(), C.c_void_p()
    h = k.b200_csr_symmetric_from_triu(n, p.ctypes.data, i.ctypes.data, x.ctypes.data, C.byref(mu), C.byref(ml))
    assert h
    full = (U + sp.triu(U, 1).T).tocsr()
    # structurally full diagonal: add explicit zeros where missing
    pattern = (full != 0).astype(np.int8) + sp.eye(n, format="csr", dtype=np.int8)
    rows, cols = pattern.nonzero()
    order = np.lexsort((cols, rows))
    rows, cols = rows[order], cols[order]
    vals = np.asarray(full[rows, cols]).ravel()
    rp_ref = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n))]).astype(np.int32)
    nnzf = k.b200_csr_nnz(h)
    assert nnzf == rows.size
    rp = np.zeros(n + 1, dtype=np.int32)
    ci = np.zeros(nnzf, dtype=np.int32)
    vx = np.zeros(nnzf, dtype=np.float64)
    assert k.b200_csr_download(h, rp.ctypes.data, ci.ctypes.data, vx.ctypes.data) == 0
    assert np.array_equal(rp, rp_ref) and np.array_equal(ci, cols) and np.array_equal(vx, vals)
    hu, hl = np.zeros(U.nnz, dtype=np.int32), np.zeros(U.nnz, dtype=np.int32)
    assert k.b200_copy_out(hu.ctypes.data, mu, U.nnz * 4) == 0 and k.b200_copy_out(hl.ctypes.data, ml, U.nnz * 4) == 0
    er, ec = U.tocoo().row, U.tocoo().col          # CSC order (sorted indices): same order as the data
    Uc = U.tocoo()
    order_csc = np.lexsort((Uc.row, Uc.col))
    er, ec, ev = Uc.row[order_csc], Uc.col[order_csc], Uc.data[order_csc]
    row_of = np.repeat(np.arange(n), np.diff(rp))
    assert np.array_equal(row_of[hu], er) and np.array_equal(ci[hu], ec) and np.array_equal(vx[hu], ev)
    off = er != ec
    assert (hl[~off] == -1).all()
    assert np.array_equal(row_of[hl[off]], ec[off]) and np.array_equal(ci[hl[off]], er[off])
    assert np.array_equal(vx[hl[off]], ev[off])
    k.b200_free(mu)
    k.b200_free(ml)
    k.b200_csr_destroy(h)
