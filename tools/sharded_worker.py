"""torchrun worker: solve one QP row-sharded over all ranks and compare with the CPU oracle
(small sizes) or report timings (large sizes).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/sharded_worker.py \
        --family lasso --scale 0.01 --check
"""
import argparse, json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from osqp_b200 import problems
from osqp_b200.dist import ShardedOSQP, init_sharded

ap = argparse.ArgumentParser()
ap.add_argument("--family", default="lasso")
ap.add_argument("--scale", type=float, default=0.01)
ap.add_argument("--eps", type=float, default=1e-3)
ap.add_argument("--check", action="store_true")
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--layout", default="split", choices=["split", "rows"])
ap.add_argument("--blocks", action="store_true",
                help="svm / huber: every rank generates only its own sample blocks (problems.*_shard) and sets "
                     "up through ShardedOSQP.setup_local -- no rank ever holds the whole problem")
args = ap.parse_args()

local = int(os.environ.get("LOCAL_RANK", "0"))
if os.environ.get("B200_TRACE_FILE"):     # one launch trace per rank
    os.environ["B200_TRACE_FILE"] += f".rank{os.environ.get('RANK', '0')}"
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
k = init_sharded(dist, local)

s = args.scale
shard = None
if args.blocks:
    gen = {"svm": problems.svm_shard, "huber": problems.huber_shard}[args.family]
    gkw = dict(n_features=int(1e4 * min(1, s * 10)), n_samples=int(1e7 * s), density=1e-3, seed=1,
               block=max(64, int(1e7 * s) // 37))
    shard = gen(rank, world, **gkw)
    pb = gen(0, 1, **gkw) if (args.check and rank == 0) else dict(P=shard["P"], A=shard["A"])
elif args.family == "lasso":
    pb = problems.lasso(int(1e5 * s), int(1e6 * s), density=min(1.0, 1e-4 / s) if s < 1 else 1e-4)
elif args.family == "portfolio":
    pb = problems.portfolio(int(1e6 * s), int(1e4 * s), density=min(0.5, 1e-2 / s) if s < 1 else 1e-2)
elif args.family == "huber":
    pb = problems.huber(int(1e4 * min(1, s * 10)), int(1e7 * s), density=1e-3)
elif args.family == "svm":
    pb = problems.svm(int(1e4 * min(1, s * 10)), int(1e7 * s), density=1e-3)
else:
    pb = problems.random_qp(int(1e4 * s * 100), int(2e4 * s * 100))
n, m = (shard["n_global"], shard["m_global"]) if shard is not None else (pb["P"].shape[0], pb["A"].shape[0])
kw = dict(eps_abs=args.eps, eps_rel=args.eps, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
          verbose=0, warm_starting=0)
if args.check:
    kw.update(eps_abs=1e-6, eps_rel=1e-6, cg_tol_fraction=1e-8, cg_max_iter=500, max_iter=20000)

prob = ShardedOSQP(rank, world, layout=args.layout)
t0 = time.perf_counter()
if shard is not None:
    prob.setup_local(shard, n, m, **kw)
else:
    prob.setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
t1 = time.perf_counter()
best = None
for rep in range(args.reps):
    dist.barrier(); torch.cuda.synchronize()
    k.b200_trace_mark(b"solve-begin")
    ta = time.perf_counter(); r = prob.solve(); tb = time.perf_counter()
    k.b200_trace_mark(b"solve-end")
    tmax = torch.tensor([tb - ta], device="cuda", dtype=torch.float64); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    best = tmax.item() if best is None else min(best, tmax.item())
cg, ns = prob.cg_stats()
if shard is not None:      # assemble through the shard's own global ids
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(r.x), np.asarray(r.y), shard["cols"], shard["rows"]))
    x, y = np.zeros(n), np.zeros(m)
    for xr, yr, cols, rows in parts:
        x[cols] = xr
        y[rows] = yr[:len(rows)]
    n_shared_out = int(shard["n_shared"])
else:
    x, y = prob.gather(r, dist)
    n_shared_out = int(prob.plan["shared"].size) if prob.plan is not None else n
k.b200_dist_p2p_enabled.restype = __import__("ctypes").c_int
k.b200_graph_launch_count.restype = __import__("ctypes").c_ulonglong
p2p_err = int(k.b200_dist_p2p_error())
ncalls = __import__("ctypes").c_ulonglong(0); nbytes = __import__("ctypes").c_ulonglong(0)
k.b200_dist_stats(__import__("ctypes").byref(ncalls), __import__("ctypes").byref(nbytes))
if rank == 0:
    out = dict(family=args.family, n=n, m=m, nnzA=int(pb["A"].nnz), world=world, p2p=bool(k.b200_dist_p2p_enabled()), graph_launches=int(k.b200_graph_launch_count()), p2p_error=p2p_err,
               blocks=bool(args.blocks), status=r.info.status, iters=r.info.iter,
               obj=r.info.obj_val, prim_res=r.info.prim_res, dual_res=r.info.dual_res, cg_iters=cg, solves=ns,
               setup_s=t1 - t0, solve_s=best, iters_per_s=r.info.iter / best, allreduce_calls=ncalls.value,
               allreduce_MB=nbytes.value / 1e6, layout=args.layout,
               n_shared=n_shared_out,
               n_local=int(prob.n), m_local=int(prob.m))
    if args.check:
        from osqp_b200.interface import OSQP as G, LoadedLibrary
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        kwo = {kk: vv for kk, vv in kw.items() if not kk.startswith("cg_")}
        ro = G(LoadedLibrary(os.path.join(root, "oracle/_ref/libosqp_builtin.so"))).setup(
            pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kwo).solve()
        out.update(oracle_status=ro.info.status, oracle_iters=ro.info.iter, oracle_obj=ro.info.obj_val,
                   obj_rel_err=abs(r.info.obj_val - ro.info.obj_val) / max(1.0, abs(ro.info.obj_val)),
                   x_err=float(np.abs(x - ro.x).max()), y_err=float(np.abs(y - ro.y).max()))
        ok = (p2p_err == 0 and out["status"] == out["oracle_status"] and out["obj_rel_err"] <= 1e-6 and
              out["x_err"] <= 1e-4 * max(1.0, np.abs(ro.x).max()) and
              abs(out["iters"] - ro.info.iter) <= max(0.1 * ro.info.iter, 50))
        out["PARITY"] = "OK" if ok else "FAIL"
    print("SHARDED " + json.dumps(out), flush=True)
# every rank must hold the same replicated entries of x (all of x under plain row sharding, the
# shared slice under the column split) -- checksum of checksums
ns = n_shared_out
rep = np.asarray(r.x)[:ns]
xs = torch.tensor([float(np.sum(rep)), float(np.abs(rep).max()) if ns else 0.0], device="cuda", dtype=torch.float64)
lo, hi = xs.clone(), xs.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
if rank == 0:
    print("REPLICATED_X_IDENTICAL", bool((lo == hi).all().item()), flush=True)
prob.cleanup()
dist.barrier()
k.b200_dist_finalize()
k.b200_shutdown()
dist.destroy_process_group()
