"""Reference algebra/cuda backend (unmodified sources, compiled for sm_100: oracle/_ref/
libosqp_refcuda_*.so) against the B200 backend on the same GPU, same problem, same settings
(BASELINE.md B2).  Each library runs in its own process (both export the same symbols)."""
import json, subprocess, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = r'''
import sys, time, json
sys.path.insert(0, %(root)r)
import numpy as np
from osqp_b200 import problems
from osqp_b200.interface import OSQP, LoadedLibrary
which, prec, scale, family = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
if family == "lasso":
    pb = problems.lasso(int(1e5 * scale), int(1e6 * scale), density=1e-4 if scale >= 1 else 1e-4 / scale)
elif family == "random_qp":
    pb = problems.random_qp(10000, 20000, 200000)
elif family == "mpc":
    pb = problems.mpc(N=12)
kw = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
          polishing=0, verbose=0, warm_starting=0, linsys_solver=2)
if prec == "f32":
    kw["check_dualgap"] = 0
dt = np.float64 if prec == "f64" else np.float32
if which == "b200":
    from osqp_b200 import load_library
    L = load_library(prec)
else:
    L = LoadedLibrary(%(root)r + "/oracle/_ref/libosqp_refcuda_" + prec + ".so", dt)
res = []
for rep in range(4):
    t0 = time.perf_counter(); s = OSQP(L).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw); t1 = time.perf_counter()
    r = s.solve(); t2 = time.perf_counter()
    r2 = s.solve(); t3 = time.perf_counter()
    res.append(dict(setup_s=t1 - t0, solve_s=t2 - t1, resolve_s=t3 - t2, iters=r.info.iter, status=r.info.status, obj=r.info.obj_val))
    s.cleanup()
best = min(res[1:], key=lambda d: d["solve_s"])
print("RESULT " + json.dumps(dict(which=which, prec=prec, family=family, n=pb["P"].shape[0], m=pb["A"].shape[0], nnzA=int(pb["A"].nnz), **best,
      iters_per_s=best["iters"] / best["solve_s"], e2e_iters_per_s=best["iters"] / (best["solve_s"] + best["setup_s"]))))
'''
if __name__ == "__main__":
    scale = sys.argv[1] if len(sys.argv) > 1 else "1.0"
    fams = sys.argv[2].split(",") if len(sys.argv) > 2 else ["lasso"]
    for fam in fams:
        for prec in ("f64", "f32"):
            for which in ("refcuda", "b200"):
                p = subprocess.run([sys.executable, "-c", WORKER % dict(root=ROOT), which, prec, scale, fam],
                                   capture_output=True, text=True, timeout=900)
                lines = [l for l in p.stdout.splitlines() if l.startswith("RESULT")]
                print(lines[0] if lines else f"FAILED {which} {prec} {fam}: " + p.stderr[-600:], flush=True)
