"""Why does configs[3] SVM solve in 326 ms from problems.svm and in 393 ms from problems.svm_shard(0, 1)?
Times six consecutive solves of each (wall clock, synchronised), prints per-pass times of the CG loop."""
import sys, time, json, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from osqp_b200 import OSQP, problems
from osqp_b200.devmem import kernels
import ctypes as C
k = kernels(); assert k.b200_init(0) == 0
KW = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5, polishing=0,
          verbose=0, warm_starting=0)
scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
ns = int(10_000_000 * scale)
GENS = {"block": ("block-seeded svm_shard(0,1)", lambda: problems.svm_shard(0, 1, 10_000, ns, 1e-3)),
        "global": ("global problems.svm", lambda: problems.svm(10_000, ns, 1e-3))}
for key in (sys.argv[2:] or ["block", "global"]):
    name, gen = GENS[key]
    pb = gen()
    A = pb["A"]
    rl = np.diff(A.tocsr().indptr)
    s = OSQP().setup(pb["P"], pb["q"], A, pb["l"], pb["u"], **KW)
    times = []
    for rep in range(6):
        k.b200_sync()
        t0 = time.perf_counter(); r = s.solve(); k.b200_sync(); times.append(round(1e3 * (time.perf_counter() - t0), 1))
    out = (C.c_double * 14)()
    k.b200_pcg_profile_last.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
    k.b200_pcg_profile_last.restype = C.c_int; k.b200_pcg_profile_last(5, out, 14)
    print(json.dumps(dict(gen=name, nnzA=int(A.nnz), iters=r.info.iter, status=r.info.status, solve_ms=times,
                          passes_us=dict(pass_A=round(out[0],1), pass_K2=round(out[1],1), update=round(out[2],1), iteration=round(out[3],1), init=round(out[5],1)), indices_dtype=str(A.indices.dtype),
                          sorted=bool(A.has_sorted_indices), row_len_max=int(rl.max()), row_len_mean=float(rl.mean()))), flush=True)
    s.cleanup(); del pb, s, A
k.b200_shutdown()
