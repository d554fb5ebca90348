"""Oracle answers for the BASELINE.json configs at sizes where the PCG *graph* driver runs.

The CPU oracle (oracle/_ref/libosqp_builtin.so = unmodified reference core + builtin backend + the
QDLDL restatement) needs minutes for some of these (configs[0] at full size: 200 s per solve, the
KKT factor of the 1e4-variable random QP is almost dense), so they are solved HERE once and the
answers are committed as baseline_<case>.npz; the GPU parity tests (tests/test_gpu_baseline_parity.py)
regenerate the same seeded problem, check its fingerprint against the fixture and compare the B200
solution with the stored oracle solution.  Nothing of the problem data is stored, only the oracle's
answers (status, iterations, objective, residuals, x, y) and a fingerprint of the inputs.

    python tests/golden/make_baseline_golden.py [case ...]      # needs oracle/_ref/libosqp_builtin.so
"""
import os
import sys
import time
from concurrent.futures import ProcessPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

# the settings bench.py uses (the three CUDA-conditional defaults pinned so that every
# backend runs the same algorithm) ...
BENCH = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
             polishing=0, verbose=0, warm_starting=0, max_iter=4000)
# ... and the tight run of test_tight_parity_with_builtin_qdldl
TIGHT = dict(eps_abs=3e-7, eps_rel=3e-7, rho_is_vec=0, check_termination=25, polishing=0, verbose=0,
             max_iter=20000)


def build(case):
    from osqp_b200 import problems
    if case == "random_qp_full":      # BASELINE configs[0] at full size
        return problems.random_qp(10_000, 20_000, 200_000, seed=1)
    if case == "lasso_s002":          # bench.py's CPU sample of configs[1]: scale 0.02
        return problems.lasso(2000, 20000, density=1e-4 / 0.02, seed=1)
    if case == "lasso_mid":           # 2.1e6 entries in A: the lean passes run
        return problems.lasso(4000, 200_000, density=2.5e-3, seed=1)
    if case == "huber_mid":
        return problems.huber(200, 100_000, density=0.05, seed=1)
    if case == "svm_mid":
        return problems.svm(200, 100_000, density=0.05, seed=1)
    if case == "portfolio_mid":
        return problems.portfolio(100_000, 200, density=0.05, seed=1)
    if case == "mpc_N12":             # one instance of BASELINE configs[4]
        return problems.mpc(N=12, seed=1)
    raise KeyError(case)


CASES = ["random_qp_full", "lasso_s002", "lasso_mid", "huber_mid", "svm_mid", "portfolio_mid", "mpc_N12"]


def fingerprint(pb):
    """order-sensitive checksums of the generated inputs (guards against a generator / numpy change)"""
    import scipy.sparse as sp
    A, P = sp.csc_matrix(pb["A"]), sp.csc_matrix(pb["P"])
    w = np.cos(np.arange(A.nnz) * 0.37)
    fin = lambda v: np.where(np.isfinite(v), v, 0.0)
    return np.array([A.shape[0], A.shape[1], A.nnz, P.nnz, float(A.data @ w), float(A.indices.astype(np.float64) @ w),
                     float(P.data.sum()), float(np.asarray(pb["q"]).sum()), float(fin(pb["l"]).sum()),
                     float(fin(pb["u"]).sum())])


MAX_STORED = 16384


def sample(tag, x, y):
    """x and y are stored as strided samples of at most MAX_STORED entries (the fixtures stay a few
    hundred KB) plus whole-vector norms; the tests apply the same stride to the B200 solution."""
    out = {}
    for nm, v in (("x", x), ("y", y)):
        stride = max(1, -(-v.size // MAX_STORED))
        out[f"{tag}_{nm}"] = v[::stride].copy()
        out[f"{tag}_{nm}_stride"] = np.int64(stride)
        out[f"{tag}_{nm}_norms"] = np.array([np.abs(v).max(initial=0.0), np.sqrt(v @ v), v.sum()])
    return out


def solve_case(case):
    from osqp_b200.interface import OSQP, LoadedLibrary
    lib = LoadedLibrary(os.path.join(ROOT, "oracle", "_ref", "libosqp_builtin.so"))
    pb = build(case)
    out = {"fingerprint": fingerprint(pb)}
    runs = [("bench", BENCH), ("tight", TIGHT)]
    if case == "portfolio_mid":
        # the duality-gap criterion of OSQP 1.0 keeps this family running to max_iter on the
        # reference itself (tight: 20000 iterations, 240 s, still open); without it: 370 iterations
        runs = [("bench", BENCH), ("nogap", dict(BENCH, check_dualgap=0))]
    for tag, st in runs:
        t0 = time.time()
        s = OSQP(lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **st)
        r = s.solve()
        dt = time.time() - t0
        out.update({f"{tag}_status": np.int64(r.info.status_val), f"{tag}_iter": np.int64(r.info.iter),
                    f"{tag}_obj": np.float64(r.info.obj_val), f"{tag}_prim_res": np.float64(r.info.prim_res),
                    f"{tag}_dual_res": np.float64(r.info.dual_res), f"{tag}_seconds": np.float64(dt),
                    f"{tag}_setup_time": np.float64(r.info.setup_time), f"{tag}_solve_time": np.float64(r.info.solve_time)})
        out.update(sample(tag, np.asarray(r.x, dtype=np.float64), np.asarray(r.y, dtype=np.float64)))
        print(f"{case:16s} {tag:5s} n={pb['P'].shape[0]} m={pb['A'].shape[0]} nnzA={pb['A'].nnz} "
              f"{r.info.status} it={r.info.iter} obj={r.info.obj_val:.9e} {dt:.1f}s", flush=True)
        s.cleanup()
    np.savez_compressed(os.path.join(HERE, f"baseline_{case}.npz"), **out)
    return case


if __name__ == "__main__":
    cases = sys.argv[1:] or CASES
    with ProcessPoolExecutor(max_workers=min(len(cases), 6)) as ex:
        for c in ex.map(solve_case, cases):
            print("wrote", c)
