"""BASELINE configs[2] (Portfolio) -- why it ends in 'maximum iterations reached'.
Runs the same seeded generator through the B200 backend and the reference's own algebra/cuda build
(oracle/_ref/libosqp_refcuda_f64.so, unmodified sources) with bench.py's settings, with and without
OSQP 1.0's duality-gap criterion, and prints one JSON row per run (-> profiles/r02_portfolio_maxiter.md).
Each library runs in its own process (both export the same symbols).

    python tools/portfolio_study.py <n_assets> <k_factors> <density> [max_iter] [which,...]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, time, json
sys.path.insert(0, %(root)r)
import numpy as np
from osqp_b200 import problems
from osqp_b200.interface import OSQP, LoadedLibrary
which, n, k, dens, max_iter = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4]), int(sys.argv[5])
t0 = time.perf_counter(); pb = problems.portfolio(n, k, density=dens, seed=1); tg = time.perf_counter() - t0
base = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
            polishing=0, verbose=0, warm_starting=0, max_iter=max_iter, linsys_solver=2)
if which == "b200":
    from osqp_b200 import load_library
    L = load_library("f64")
else:
    L = LoadedLibrary(%(root)r + "/oracle/_ref/libosqp_refcuda_f64.so", np.float64)
for tag, kw in (("bench", {}), ("no_dualgap", dict(check_dualgap=0)), ("no_dualgap_cg200", dict(check_dualgap=0, cg_max_iter=200))):
    t0 = time.perf_counter(); s = OSQP(L).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **dict(base, **kw)); t1 = time.perf_counter()
    r = s.solve(); t2 = time.perf_counter()
    row = dict(which=which, run=tag, n=pb["P"].shape[0], m=pb["A"].shape[0], nnzA=int(pb["A"].nnz), status=r.info.status,
               iters=r.info.iter, obj=r.info.obj_val, prim_res=r.info.prim_res, dual_res=r.info.dual_res,
               duality_gap=r.info.duality_gap, rho_updates=r.info.rho_updates, setup_s=round(t1 - t0, 3),
               solve_s=round(t2 - t1, 3), gen_s=round(tg, 1))
    if which == "b200":
        import ctypes as C
        it, ns = C.c_longlong(0), C.c_longlong(0)
        L.lib.osqp_b200_cg_stats(C.cast(s._solver, C.c_void_p), C.byref(it), C.byref(ns))
        row["cg_per_admm"] = round(it.value / max(ns.value, 1), 2)
    print("ROW " + json.dumps(row), flush=True)
    s.cleanup()
'''
if __name__ == "__main__":
    n, k, dens = sys.argv[1], sys.argv[2], sys.argv[3]
    max_iter = sys.argv[4] if len(sys.argv) > 4 else "4000"
    whiches = sys.argv[5].split(",") if len(sys.argv) > 5 else ["b200", "refcuda"]
    for which in whiches:
        p = subprocess.run([sys.executable, "-c", WORKER % dict(root=ROOT), which, n, k, dens, max_iter],
                           capture_output=True, text=True, timeout=1500)
        rows = [l for l in p.stdout.splitlines() if l.startswith("ROW")]
        print("\n".join(rows) if rows else f"FAILED {which}: " + p.stderr[-800:], flush=True)
