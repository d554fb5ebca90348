"""All five BASELINE.json configs on one B200: setup / solve time, ADMM and CG iterations, ADMM it/s.
Writes a markdown table (development aid; results copied to profiles/)."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels
k = kernels(); assert k.b200_init(0) == 0
KW = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5, verbose=0, warm_starting=0)
cases = [
    ("configs[0] Random QP n=1e4 m=2e4", lambda: problems.random_qp(10_000, 20_000, 200_000)),
    ("configs[1] Lasso 1e5 x 1e6", lambda: problems.lasso(100_000, 1_000_000, 1e-4)),
    ("configs[2] Portfolio k=1e4 n=1e6 (nnz F 1e8)", lambda: problems.portfolio(1_000_000, 10_000, 1e-2)),
    ("configs[3] Huber 1e7 x 1e4 (1e8 nnz)", lambda: problems.huber(10_000, 10_000_000, 1e-3)),
    ("configs[3] SVM 1e7 x 1e4 (1e8 nnz)", lambda: problems.svm(10_000, 10_000_000, 1e-3)),
    ("configs[4] one MPC QP (N=12)", lambda: problems.mpc(N=12)),
]
only = sys.argv[1:] 
rows = []
for name, gen in cases:
    if only and not any(o in name for o in only): continue
    t0 = time.perf_counter(); pb = gen(); tg = time.perf_counter() - t0
    n, m = pb["P"].shape[0], pb["A"].shape[0]
    t0 = time.perf_counter(); s = OSQP().setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **KW); ts = time.perf_counter() - t0
    best = None
    for rep in range(3):
        t0 = time.perf_counter(); r = s.solve(); dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    cg, ns = s.cg_stats()
    row = dict(config=name, n=n, m=m, nnzA=int(pb["A"].nnz), status=r.info.status, admm_iters=r.info.iter,
               cg_per_admm=round(cg / max(ns, 1), 2), setup_ms=round(1e3 * ts, 1), solve_ms=round(1e3 * best, 2),
               admm_it_per_s=round(r.info.iter / best, 1), obj=r.info.obj_val, gen_s=round(tg, 1))
    print("ROW " + json.dumps(row), flush=True)
    rows.append(row)
    s.cleanup()
    del pb, s
print("| config | n | m | nnz(A) | status | ADMM it | CG / ADMM | setup ms | solve ms | ADMM it/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for r in rows:
    print(f"| {r['config']} | {r['n']} | {r['m']} | {r['nnzA']} | {r['status']} | {r['admm_iters']} | {r['cg_per_admm']} | {r['setup_ms']} | {r['solve_ms']} | {r['admm_it_per_s']} |")
