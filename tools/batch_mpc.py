"""BASELINE configs[4]: a batch of independent MPC QPs (docs/examples/mpc.rst:30-90 quadcopter, N = 12:
n = 204, m = 360) that differ in the initial state, through the batched one-CTA-per-QP kernel
(OSQP.solve_batch -> osqp_b200/csrc/batch.cu), with the CPU oracle's QPs/s on one host core beside it.

    python tools/batch_mpc.py [nb] [--cpu-sample K]        # prints one JSON line
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osqp_b200 import OSQP, problems

SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
                polishing=0, verbose=0, warm_starting=0)


def mpc_batch(nb, seed=1):
    rng = np.random.default_rng(seed)
    base = problems.mpc(N=12, seed=1)
    nx = base["nx"]
    x0s = 0.1 * (2 * rng.random((nb, nx)) - 1)
    L, U = np.tile(base["l"], (nb, 1)), np.tile(base["u"], (nb, 1))
    L[:, :nx] = -x0s
    U[:, :nx] = -x0s
    return base, L, U


def run_batch(nb, reps=5, prec="f64", settings=SETTINGS):
    import ctypes as C
    base, L, U = mpc_batch(nb)
    t0 = time.perf_counter()
    tmpl = OSQP(prec).setup(base["P"], base["q"], base["A"], base["l"], base["u"], **settings)
    t_setup = time.perf_counter() - t0
    tmpl._lib.osqp_b200_last_batch_kernel_ms.restype = C.c_double
    tmpl.solve_batch(L, U)
    best, best_k = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        r = tmpl.solve_batch(L, U)
        dt = time.perf_counter() - t0
        km = tmpl._lib.osqp_b200_last_batch_kernel_ms()
        if best is None or dt < best:
            best, best_k = dt, km
    tmpl.cleanup()
    return dict(nb=nb, n=int(base["P"].shape[0]), m=int(base["A"].shape[0]), setup_ms=1e3 * t_setup,
                e2e_ms=1e3 * best, kernel_ms=best_k, qps_per_s_e2e=nb / best, qps_per_s_kernel=nb / (best_k / 1e3),
                admm_iters_total=int(r.iter.sum()), admm_it_per_s_kernel=float(r.iter.sum()) / (best_k / 1e3),
                admm_it_per_s_e2e=float(r.iter.sum()) / best, iters_mean=float(r.iter.mean()),
                cg_per_admm=float(r.cg_iters.sum()) / float(r.iter.sum()), solved=int((r.status_val == 1).sum()),
                h2d_bytes=int(L.nbytes + U.nbytes), d2h_bytes=int(r.x.nbytes + r.y.nbytes + 7 * 4 * nb)), r, (base, L, U)


def run_cpu(base, L, U, k, settings=SETTINGS):
    """the oracle on one host core: one setup, then osqp_update_data_vec + osqp_solve per QP (the
    production pattern of docs/examples/mpc.rst:91-105)"""
    from osqp_b200.interface import OSQP as G, LoadedLibrary
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = LoadedLibrary(os.path.join(root, "oracle", "_ref", "libosqp_builtin.so"))
    s = G(lib).setup(base["P"], base["q"], base["A"], base["l"], base["u"], **settings)
    t0 = time.perf_counter()
    its, objs = 0, []
    for i in range(k):
        s.update(l=L[i], u=U[i])
        r = s.solve()
        its += r.info.iter
        objs.append(r.info.obj_val)
    dt = time.perf_counter() - t0
    s.cleanup()
    return dict(qps=k, seconds=dt, qps_per_s=k / dt, admm_it_per_s=its / dt, iters_mean=its / k), np.array(objs)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("nb", type=int, nargs="?", default=4096)
    ap.add_argument("--cpu-sample", type=int, default=256)
    args = ap.parse_args()
    from osqp_b200.devmem import kernels
    assert kernels("f64").b200_init(0) == 0
    gpu, r, (base, L, U) = run_batch(args.nb)
    out = {"batch": gpu}
    if args.cpu_sample:
        cpu, objs = run_cpu(base, L, U, min(args.cpu_sample, args.nb))
        out["cpu_oracle_1_core"] = cpu
        out["obj_rel_diff_max"] = float(np.max(np.abs(r.obj_val[:len(objs)] - objs) / np.maximum(1.0, np.abs(objs))))
        out["speedup_qps_e2e"] = gpu["qps_per_s_e2e"] / cpu["qps_per_s"]
    print("BATCH " + json.dumps(out))
