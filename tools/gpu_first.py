"""First GPU bring-up script (development aid): SpMV vs scipy, basic QP vs oracle, a random QP."""
import sys, time
import numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels, DeviceArray, csr_to_device
from osqp_b200.interface import OSQP as GenericOSQP, LoadedLibrary

k = kernels()
assert k.b200_init(0) == 0
buf = (b" " * 256)
import ctypes as C
name = C.create_string_buffer(256); k.b200_device_name(name, 256); print("device:", name.value)

rng = np.random.default_rng(0)
# --- vector reductions
v = rng.standard_normal(100003)
dv = DeviceArray(k, v)
print("norm_inf", k.b200_vec_norm_inf(dv.ptr, v.size), np.abs(v).max())
print("dot", k.b200_vec_dot(dv.ptr, dv.ptr, v.size), v @ v)
# --- SpMV incl. long rows
for (m, n, dens) in [(1000, 800, 0.01), (50, 30000, 0.5), (20000, 7, 0.3), (3, 3, 0.0)]:
    M = sp.random(m, n, density=dens, format="csr", random_state=1)
    if m == 50:
        M = sp.vstack([M, sp.csr_matrix(np.ones((1, n)))]).tocsr()
    h = csr_to_device(k, M)
    x = rng.standard_normal(M.shape[1]); y = rng.standard_normal(M.shape[0])
    dx, dy = DeviceArray(k, x), DeviceArray(k, y)
    k.b200_csr_spmv(h, dx.ptr, dy.ptr, 2.0, -0.5)
    ref = 2.0 * (M @ x) - 0.5 * y
    err = np.abs(dy.get() - ref).max() if ref.size else 0.0
    k.b200_csr_spmv(h, dx.ptr, dy.ptr, 1.0, 0.0)
    err2 = np.abs(dy.get() - M @ x).max() if ref.size else 0.0
    out = DeviceArray(k, n=M.shape[0]); k.b200_csr_row_absmax(h, out.ptr)
    err3 = np.abs(out.get() - abs(M).max(axis=1).toarray().ravel()).max()
    print("spmv", M.shape, M.nnz, "err", err, err2, err3)
    k.b200_csr_destroy(h)

# --- basic QP through the full OSQP stack
P = sp.triu(sp.csc_matrix([[4., 1.], [1., 2.]]), format='csc'); q = np.ones(2)
A = sp.csc_matrix(np.array([[1., 1.], [1., 0.], [0., 1.], [0., 1.]]))
l = np.array([1., 0., 0., -np.inf]); u = np.array([1., 0.7, 0.7, np.inf])
s = OSQP().setup(P, q, A, l, u, rho=0.1, alpha=1.6, max_iter=2000, scaling=1, eps_abs=1e-5, eps_rel=1e-5, verbose=1)
r = s.solve()
print("basic_qp", r.x, r.y, r.info.obj_val, r.info.status, r.info.iter, "cg", s.cg_stats())

oracle = LoadedLibrary("oracle/_ref/libosqp_builtin.so")
for (n, m) in [(200, 400), (2000, 4000), (10000, 20000)]:
    pb = problems.random_qp(n, m, nnz_target=20 * n)
    kw = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5)
    t0 = time.time(); s = OSQP().setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw); t1 = time.time()
    r = s.solve(); t2 = time.time()
    o = GenericOSQP(oracle).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw); ro = o.solve()
    print(f"random_qp n={n}: b200 {r.info.status} it={r.info.iter} obj={r.info.obj_val:.6e} setup={t1-t0:.3f}s solve={t2-t1:.3f}s cg={s.cg_stats()}"
          f" | oracle {ro.info.status} it={ro.info.iter} obj={ro.info.obj_val:.6e} solve={ro.info.solve_time:.3f}s")
    s.cleanup(); o.cleanup()
print("last_error", k.b200_last_error(), "launches", k.b200_launch_count())
