"""Development probe: raw kernel bandwidth + full solve timing on one workload."""
import sys, time, argparse
import numpy as np, scipy.sparse as sp
sys.path.insert(0, ".")
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels, DeviceArray, csr_to_device

ap = argparse.ArgumentParser()
ap.add_argument("--family", default="lasso")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--solve", type=int, default=1)
ap.add_argument("--eps", type=float, default=1e-3)
args = ap.parse_args()
k = kernels(); assert k.b200_init(0) == 0
F = 8
t0 = time.time()
if args.family == "lasso":
    pb = problems.lasso(int(1e5 * args.scale), int(1e6 * args.scale), density=1e-4 / args.scale if args.scale < 1 else 1e-4)
elif args.family == "portfolio":
    pb = problems.portfolio(int(1e6 * args.scale), int(1e4 * args.scale))
elif args.family == "huber":
    pb = problems.huber(int(1e4), int(1e7 * args.scale))
elif args.family == "svm":
    pb = problems.svm(int(1e4), int(1e7 * args.scale))
else:
    pb = problems.random_qp()
A = pb["A"].tocsr(); At = pb["A"].T.tocsr(); n = A.shape[1]; m = A.shape[0]
print(f"gen {time.time()-t0:.1f}s  n={n} m={m} nnzA={A.nnz} nnzP={pb['P'].nnz}", flush=True)

def time_kernel(fn, reps=20):
    e0, e1 = k.b200_event_create(), k.b200_event_create()
    for _ in range(3): fn()
    k.b200_event_record(e0)
    for _ in range(reps): fn()
    k.b200_event_record(e1)
    return k.b200_event_elapsed_ms(e0, e1) / reps

rng = np.random.default_rng(0)
for name, M in (("A", A), ("At", At)):
    h = csr_to_device(k, M)
    x = DeviceArray(k, rng.standard_normal(M.shape[1])); y = DeviceArray(k, n=M.shape[0])
    ms = time_kernel(lambda: k.b200_csr_spmv(h, x.ptr, y.ptr, 1.0, 0.0))
    byts = M.nnz * (F + 4) + (M.shape[0] + 1) * 4 + M.shape[1] * F + M.shape[0] * F
    print(f"spmv {name}: {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s  ({byts/1e6:.1f} MB)", flush=True)
    k.b200_csr_destroy(h)
# banded matrix with the same row lengths as A but consecutive columns: cost of everything
# except the random gather
Ab = A.copy(); Ab.sort_indices()
rl = np.diff(Ab.indptr); starts = (np.arange(m) * 7) % max(1, n - rl.max() - 1)
Ab.indices = (np.repeat(starts, rl) + (np.arange(Ab.nnz) - np.repeat(Ab.indptr[:-1], rl))).astype(np.int32)
h = csr_to_device(k, Ab); x = DeviceArray(k, rng.standard_normal(n)); y = DeviceArray(k, n=m)
ms = time_kernel(lambda: k.b200_csr_spmv(h, x.ptr, y.ptr, 1.0, 0.0))
byts = Ab.nnz * (F + 4) + (m + 1) * 4 + n * F + m * F
print(f"spmv banded(A): {ms*1e3:.1f} us  {byts/ms/1e6:.0f} GB/s", flush=True)
k.b200_csr_destroy(h)
v1 = DeviceArray(k, rng.standard_normal(n)); v2 = DeviceArray(k, rng.standard_normal(n)); v3 = DeviceArray(k, n=n)
ms = time_kernel(lambda: k.b200_vec_add_scaled(v3.ptr, 1.0, v1.ptr, 2.0, v2.ptr, n))
print(f"add_scaled n={n}: {ms*1e3:.1f} us {3*n*F/ms/1e6:.0f} GB/s", flush=True)
big = 1 << 27
b1 = DeviceArray(k, n=big); b2 = DeviceArray(k, n=big); b3 = DeviceArray(k, n=big)
k.b200_vec_set_scalar(b1.ptr, 1.0, big); k.b200_vec_set_scalar(b2.ptr, 1.0, big)
ms = time_kernel(lambda: k.b200_vec_add_scaled(b3.ptr, 1.0, b1.ptr, 2.0, b2.ptr, big), reps=5)
print(f"add_scaled n=2^27: {ms*1e3:.1f} us {3*big*F/ms/1e6:.0f} GB/s", flush=True)
ms = time_kernel(lambda: k.b200_vec_norm_inf(b1.ptr, big), reps=5)
print(f"norm_inf n=2^27: {ms*1e3:.1f} us {big*F/ms/1e6:.0f} GB/s", flush=True)
del b1, b2, b3

# --- PCG kernel micro-benchmark: fixed number of CG iterations per launch
Pfull = sp.csr_matrix(sp.triu(pb["P"]) + sp.triu(pb["P"], 1).T + sp.eye(n) * 0)  # structurally full diag below
Pfull = (Pfull + sp.eye(n, format="csr") * 1e-300).tocsr()
hP, hA, hAt = csr_to_device(k, Pfull), csr_to_device(k, A), csr_to_device(k, At)
pcg = k.b200_pcg_create(hP, hA, hAt, n, m)
k.b200_pcg_configure(pcg, 1e-6, 0.1, None, 1, 0)
k.b200_pcg_refresh_matrices(pcg); k.b200_pcg_refresh_precond(pcg)
bvec = DeviceArray(k, rng.standard_normal(n + m)); b0 = DeviceArray(k, bvec.get()); b1v = DeviceArray(k, rng.standard_normal(n + m))
import ctypes as C
def run(K):
    e0, e1 = k.b200_event_create(), k.b200_event_create()
    tot = 0.0
    for rep in range(6):
        k.b200_copy_in(bvec.ptr, (b0 if rep % 2 else b1v).ptr, (n + m) * F)
        k.b200_event_record(e0)
        k.b200_pcg_solve(pcg, bvec.ptr, 2, 0.0, 0.0, K, 0.15, 10)
        k.b200_event_record(e1)
        ms = k.b200_event_elapsed_ms(e0, e1)
        if rep >= 2: tot += ms
    li = C.c_int(0); k.b200_pcg_stats(pcg, None, None, C.byref(li), None, None)
    return tot / 4, li.value
t0_, i0 = run(0); t10, i10 = run(10)
nnzPf = Pfull.nnz
bpi = problems.kkt_bytes_per_cg_iter(n, m, A.nnz, nnzPf)
per = (t10 - t0_) / max(i10, 1)
print(f"pcg kernel: K=0 {t0_*1e3:.1f} us (iters {i0}); K=10 {t10*1e3:.1f} us (iters {i10}); per CG iter {per*1e3:.1f} us -> {bpi/per/1e6:.0f} GB/s of {bpi/1e6:.1f} MB", flush=True)
k.b200_pcg_destroy(pcg)
for h in (hP, hA, hAt): k.b200_csr_destroy(h)

if args.solve:
    kw = dict(eps_abs=args.eps, eps_rel=args.eps, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5, verbose=1)
    t0 = time.time(); s = OSQP().setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw); t1 = time.time()
    r = s.solve(); t2 = time.time()
    cg, ns = s.cg_stats()
    nnzPf = problems.nnz_P_full(pb["P"])
    bpi = problems.kkt_bytes_per_cg_iter(n, m, A.nnz, nnzPf)
    print(f"solve: {r.info.status} iters={r.info.iter} cg={cg} solves={ns} setup={t1-t0:.3f}s solve={t2-t1:.3f}s "
          f"({(t2-t1)/r.info.iter*1e3:.3f} ms/iter) obj={r.info.obj_val:.6e}")
    print(f"cg-iter bytes {bpi/1e6:.1f} MB; if all time were CG: {bpi*cg/(t2-t1)/1e9:.0f} GB/s")
print("last_error", k.b200_last_error())
