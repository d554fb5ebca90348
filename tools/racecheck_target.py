"""Small workloads for compute-sanitizer racecheck (the full tests run for tens of minutes under it):
(1) the lean passes of the graph PCG driver on a matrix with rows of 2500 and 1500 entries (chunk path,
multi- and single-chunk) next to short rows, loop body as plain launches (B200_PCG_HOSTLOOP=1 -- the
sanitizers cannot follow kernel nodes of a conditional graph); (2) the batched one-CTA-per-QP kernel on 3 QPs.

    B200_PCG_HOSTLOOP=1 compute-sanitizer --tool racecheck python tools/racecheck_target.py
"""
import os, sys
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as sla
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from osqp_b200.devmem import kernels, DeviceArray
import test_gpu_kernels as tk

k = kernels("f64")
assert k.b200_init(0) == 0
rng = np.random.default_rng(0)
n, m = 3000, 12
rows = []
for i in range(m):
    cnt = [2500, 1500, 40][i % 3]
    cols = np.sort(rng.choice(n, cnt, replace=False))
    rows.append(sp.csr_matrix((rng.standard_normal(cnt), (np.zeros(cnt, dtype=int), cols)), shape=(1, n)))
A = sp.vstack(rows, format="csr")
P = (sp.eye(n) * 0.5).tocsc()
s, keep = tk._pcg(k, P, A, 1e-3, 0.7, None, 0, "graph")
rhs = rng.standard_normal(n + m)
b = DeviceArray(k, rhs)
for it in range(2):
    k.b200_copy_in(b.ptr, rhs.ctypes.data, (n + m) * 8)
    assert k.b200_pcg_solve(s, b.ptr, 2, 1e-9, 1e-9, 8, 0.15, 10) == 0
out = b.get()
K = (P + 1e-3 * sp.eye(n) + 0.7 * (A.T @ A)).tocsc()
x_ref = sla.spsolve(K, rhs[:n] + 0.7 * (A.T @ rhs[n:]))
print("lean passes with chunked rows: max |x - x_ref| =", float(np.abs(out[:n] - x_ref).max()))
k.b200_pcg_destroy(s)

from osqp_b200 import OSQP
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import batch_mpc
base, L, U = batch_mpc.mpc_batch(3)
tmpl = OSQP("f64").setup(base["P"], base["q"], base["A"], base["l"], base["u"], **dict(batch_mpc.SETTINGS, max_iter=60))
r = tmpl.solve_batch(L, U)
print("batch kernel: iters", r.iter.tolist(), "status", r.status_val.tolist())
tmpl.cleanup()
print("racecheck target done, last error", k.b200_last_error())
