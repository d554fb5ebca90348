"""Batch of independent MPC QPs on ONE GPU (BASELINE configs[4]): T host threads, each with its own
library context / stream, solve their share of the batch one after the other; the solves of
different threads overlap on the device.  Prints QPs/s and aggregate ADMM iterations/s per T."""
import argparse
import json
import sys
import threading
import time

sys.path.insert(0, ".")
import numpy as np
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels

SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
                verbose=0, warm_starting=0)
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=512)
ap.add_argument("--threads", default="1,4,8,16")
args = ap.parse_args()
k = kernels("f64")
assert k.b200_init(0) == 0
base = problems.mpc(N=12, seed=1)
rng = np.random.default_rng(1)
x0s = 0.1 * (2 * rng.random((args.batch, base["nx"])) - 1)


def run_share(idx, out):
    """Parametric re-solve: one setup per thread, new (l, u) per QP through osqp_update_data_vec."""
    s = OSQP("f64").setup(base["P"], base["q"], base["A"], base["l"], base["u"], **SETTINGS)
    nx = base["nx"]
    it = 0
    for i in idx:
        l, u = base["l"].copy(), base["u"].copy()
        l[:nx] = u[:nx] = -x0s[i]
        s.update(l=l, u=u)
        r = s.solve()
        assert r.info.status == "solved", r.info.status
        it += r.info.iter
    s.cleanup()
    out.append(it)


rows = []
for T in [int(t) for t in args.threads.split(",")]:
    out = []
    shares = [list(range(args.batch))[i::T] for i in range(T)]
    ths = [threading.Thread(target=run_share, args=(sh, out)) for sh in shares]
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    rows.append(dict(threads=T, batch=args.batch, seconds=round(dt, 3), qps_per_s=round(args.batch / dt, 1),
                     admm_it_per_s=round(sum(out) / dt, 1), admm_it_per_qp=round(sum(out) / args.batch, 1)))
    print("BATCH " + json.dumps(rows[-1]), flush=True)
k.b200_shutdown()
