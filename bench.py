#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 OSQP backend (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

metric : ADMM iterations / second (whole job) at eps_abs = eps_rel = 1e-3, plus time-to-solution.
workload (N = 1): BASELINE.json configs[1] -- Lasso, 1e5 features x 1e6 samples (~1.14e7 nnz,
          n = m = 1.2e6), single B200, indirect (PCG) solver, f64.
step   : one complete osqp_solve of the workload from a cold start.
value  : iterations / s with the problem resident in HBM (solver set up before the timed region).
e2e    : the same metric through the public API from HOST arrays: every step is osqp_setup
         (host -> device copies of P, A, q, l, u and all format conversions) + osqp_solve + the
         device -> host read of the solution, on the host clock.
N > 1  : the path shards as independent QPs (BASELINE configs[4] style): every rank sets up and solves
         its own instance with no data-path collective -> "scaling": "weak".
         `--mode sharded` instead solves ONE QP with the rows of A split over the ranks (strong).
--impl reference : the reference's own CPU path (unmodified core + builtin backend + QDLDL
         restatement = oracle/_ref/libosqp_builtin.so) on a bounded sample of the same generator.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0,
                check_termination=5, polishing=0, verbose=0, warm_starting=0)
CPU_SAMPLE_SCALE = 0.02          # QDLDL fill grows ~ scale^3: 0.02 -> ~6 s, 0.04 -> ~48 s per solve
F = 8


def make_problem(scale, seed, workload="lasso"):
    """BASELINE.json workloads; `scale` shrinks them proportionally (1.0 = the named size)."""
    from osqp_b200 import problems
    if workload == "lasso":        # configs[1]
        if scale >= 1.0:
            return problems.lasso(int(1e5 * scale), int(1e6 * scale), density=1e-4, seed=seed)
        return problems.lasso(int(1e5 * scale), int(1e6 * scale), density=1e-4 / scale, seed=seed)
    if workload == "portfolio":    # configs[2]: k = 1e4 factors, n = 1e6 assets, nnz(F) = 1e8
        return problems.portfolio(int(1e6 * scale), int(1e4 * min(1.0, scale * 10) if scale < 1 else 1e4),
                                  density=1e-2, seed=seed)
    if workload == "huber":        # configs[3]
        return problems.huber(10_000, int(1e7 * scale), density=1e-3, seed=seed)
    if workload == "svm":          # configs[3]
        return problems.svm(10_000, int(1e7 * scale), density=1e-3, seed=seed)
    if workload == "random_qp":    # configs[0]
        return problems.random_qp(10_000, 20_000, 200_000, seed=seed)
    if workload == "mpc":          # configs[4], one instance
        return problems.mpc(N=12, seed=seed)
    raise ValueError(workload)


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(device), "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        rows = [r.split(",") for r in Path(self.f.name).read_text().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if "Active" in v and "Not" not in v:
                        reasons.add(nm)
            except (ValueError, IndexError):
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------ reference (CPU) arm
def oracle_library():
    from osqp_b200.interface import LoadedLibrary
    lib = ROOT / "oracle" / "_ref" / "libosqp_builtin.so"
    if not lib.exists():
        if Path("/root/reference").exists():
            subprocess.run(["make", "-C", str(ROOT / "oracle")], check=True, stdout=subprocess.DEVNULL)
        else:
            raise FileNotFoundError(f"{lib} missing and no reference tree to build it from")
    return LoadedLibrary(lib)


def cpu_step(lib, pb):
    """one end-to-end step on the CPU path: osqp_setup + osqp_solve; returns (iters, seconds,
    solve-only seconds)."""
    from osqp_b200.interface import OSQP
    t0 = time.perf_counter()
    s = OSQP(lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **SETTINGS)
    r = s.solve()
    t1 = time.perf_counter()
    it, st = r.info.iter, r.info.solve_time
    status = r.info.status
    s.cleanup()
    return it, t1 - t0, st, status


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    lib = oracle_library()
    pb = make_problem(CPU_SAMPLE_SCALE, seed=1)
    n, m = pb["P"].shape[0], pb["A"].shape[0]
    for _ in range(args.warmup):
        cpu_step(lib, pb)
    t0 = time.perf_counter()
    iters = 0
    solve_s = 0.0
    for _ in range(args.steps):
        it, dt, st, status = cpu_step(lib, pb)
        iters += it
        solve_s += st
    total = time.perf_counter() - t0
    v = iters / total
    sample = (f"Lasso generator at scale {CPU_SAMPLE_SCALE}: n={n}, m={m}, nnz(A)={pb['A'].nnz}; "
              "the full-size KKT factor does not fit QDLDL (fill ~ n_features^2)")
    out = {
        "impl": "reference", "metric": "admm_iters_per_sec", "value": v, "unit": "iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "lasso_1e5x1e6 (BASELINE configs[1]), bounded CPU sample",
                   "sample_scale": CPU_SAMPLE_SCALE, "n": n, "m": m, "nnzA": int(pb["A"].nnz),
                   "eps": 1e-3, "solver": "builtin + QDLDL (direct)", "step": "osqp_setup + osqp_solve"},
        "cpu_baseline": {"value": v, "unit": "iter/s", "cores": 1, "kind": "reference", "sample": sample,
                         "solve_only_iters_per_sec": iters / solve_s if solve_s else None},
        "e2e": {"value": v, "unit": "iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "time_to_solution_ms": 1e3 * total / args.steps, "gpu_launches": 0, "status": status,
    }
    args.emit(json.dumps(out))


# ------------------------------------------------------------------------------- B200 arm
def pcg_roofline(k, pb, kcg):
    """Time the dominant kernel (the persistent PCG kernel, one launch = one linear solve with
    exactly `kcg` CG iterations) with CUDA events on the library stream; algorithmic bytes per
    launch from SURVEY.md section 8(d)."""
    import ctypes as C
    import scipy.sparse as sp
    from osqp_b200.devmem import DeviceArray, csr_to_device
    A = pb["A"].tocsr()
    At = pb["A"].T.tocsr()
    n, m = A.shape[1], A.shape[0]
    Pu = sp.triu(pb["P"], format="csr")
    Pfull = (Pu + sp.triu(Pu, 1).T + sp.eye(n, format="csr") * 1e-300).tocsr()
    hP, hA, hAt = csr_to_device(k, Pfull), csr_to_device(k, A), csr_to_device(k, At)
    pcg = k.b200_pcg_create(hP, hA, hAt, n, m)
    k.b200_pcg_configure(pcg, 1e-6, 0.1, None, 1, 0)
    k.b200_pcg_refresh_matrices(pcg)
    k.b200_pcg_refresh_precond(pcg)
    rng = np.random.default_rng(0)
    # alternate between two right-hand sides so that every launch (warm-started from the previous
    # solution, carried A x valid: the steady state of an ADMM run) needs all `kcg` iterations
    rhs = [DeviceArray(k, rng.standard_normal(n + m)) for _ in range(2)]
    b = DeviceArray(k, n=n + m)
    e0, e1 = k.b200_event_create(), k.b200_event_create()
    times = []
    for rep in range(10):
        k.b200_copy_in(b.ptr, rhs[rep % 2].ptr, (n + m) * F)
        k.b200_event_record(e0)
        k.b200_pcg_solve(pcg, b.ptr, 2, 0.0, 0.0, kcg, 0.15, 10)
        k.b200_event_record(e1)
        ms = k.b200_event_elapsed_ms(e0, e1)
        if rep >= 4:
            times.append(ms)
    li = C.c_int(0)
    k.b200_pcg_stats(pcg, None, None, C.byref(li), None, None)
    # graph driver (the default at this size): per-kernel CUDA-event times of one CG iteration
    k.b200_pcg_profile_last.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
    k.b200_pcg_profile_last.restype = C.c_int
    ph = (C.c_double * 14)()
    phases = None
    if k.b200_pcg_profile_last(20, ph, 14) == 0:
        phases = {"pass_A_us": ph[0], "pass_K2_us": ph[1], "update_us": ph[2], "iteration_us": ph[3],
                  "initial_residual_pass_us": ph[5]}
    k.b200_pcg_destroy(pcg)
    for h in (hP, hA, hAt):
        k.b200_csr_destroy(h)
    nnzA, nnzK = A.nnz, Pfull.nnz + A.nnz

    def spmv(r, c, nnz):
        return nnz * (F + 4) + (r + 1) * 4 + c * F + r * F
    # algorithmic bytes of what the kernel must stream (osqp_b200/csrc/pcg.cu header):
    # per CG iteration SpMV(A) + SpMV([P+sigma I | A']) + 8nF; per launch one more pass over the
    # fused operator (initial residual) + (3n+3m)F of right-hand side / write-back vectors
    per_iter = spmv(m, n, nnzA) + spmv(n, n + m, nnzK) + 8 * n * F
    fixed = spmv(n, n + m, nnzK) + (3 * n + 3 * m) * F
    byts = fixed + li.value * per_iter
    ms = float(np.mean(times))
    out = {"kernel": "pcg_kernel", "cg_iters_per_launch": li.value, "bytes_per_launch": byts,
           "bytes_per_cg_iter": per_iter, "ms_per_launch": ms, "gbs": byts / ms / 1e6}
    if phases is not None:
        # the dominant kernel of the graph driver is the fused-operator pass (one launch per CG
        # iteration + one per solve): its own algorithmic bytes = SpMV([P+sigma I | A']) + the
        # p, r, M^-1 reads of the three fused dot products
        kb = spmv(n, n + m, nnzK) + 3 * n * F
        out.update({"kernel": "g_lean_pass<1> (fused-operator pass Kp = [P+sigma I | A'][p; t] + 3 dots)",
                    "bytes_per_launch": kb, "ms_per_launch": phases["pass_K2_us"] / 1e3,
                    "gbs": kb / phases["pass_K2_us"] / 1e3,
                    "solve": {"cg_iters": li.value, "bytes": byts, "ms": ms, "gbs": byts / ms / 1e6},
                    "pass_A": {"bytes": spmv(m, n, nnzA), "us": phases["pass_A_us"],
                               "gbs": spmv(m, n, nnzA) / phases["pass_A_us"] / 1e3},
                    "phases_us": phases})
    return out


def run_b200(args):
    rank, world, local = dist_env()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from osqp_b200 import OSQP, problems
    from osqp_b200.devmem import kernels
    prec = args.dtype
    sharded = args.mode == "sharded" and world > 1
    if sharded:
        # ONE QP, rows of A split over the ranks, one all-reduce of the length-n partial per K.p
        from osqp_b200.dist import ShardedOSQP, init_sharded
        k = init_sharded(dist, local, prec)
        kernels(prec)
        pb = make_problem(args.scale, seed=1, workload=args.workload)
        _Base = OSQP

        def OSQP(_prec):  # noqa: N802 -- same constructor signature as the single-GPU class
            return ShardedOSQP(rank, world, _prec)
    else:
        k = kernels(prec)
        if k.b200_init(local) != 0:
            raise RuntimeError("no usable GPU: the B200 backend has no CPU fallback")
        # same generator and seed on every rank: per-GPU work is identical, which is what "weak"
        # scaling compares (each rank still builds, uploads and solves its own copy)
        pb = make_problem(args.scale, seed=1, workload=args.workload)
    n, m = pb["P"].shape[0], pb["A"].shape[0]
    nnzA, nnzP = int(pb["A"].nnz), int(pb["P"].nnz)

    def barrier():
        k.b200_sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident-in-HBM arm: set up once, time K cold-start solves with CUDA events
    solver = OSQP(prec).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **SETTINGS)
    for _ in range(args.warmup):
        r = solver.solve()
    cg0, ns0 = solver.cg_stats()
    e0, e1 = k.b200_event_create(), k.b200_event_create()
    sampler = ClockSampler(local) if rank == 0 else None
    barrier()
    l0 = k.b200_launch_count()
    k.b200_graph_launch_count.restype = ctypes.c_ulonglong
    g0 = k.b200_graph_launch_count()
    k.b200_event_record(e0)
    iters = 0
    for _ in range(args.steps):
        r = solver.solve()
        iters += r.info.iter
    k.b200_event_record(e1)
    ms = k.b200_event_elapsed_ms(e0, e1)
    launches = k.b200_launch_count() - l0
    barrier()
    clocks = sampler.stop() if sampler else None
    cg1, ns1 = solver.cg_stats()
    # a graph launch is one enqueue but 1 + 3 k kernels (loop-init node, then A pass / operator
    # pass / fused update per CG iteration): count the kernels, not the enqueues
    k.b200_graph_launch_count.restype = ctypes.c_ulonglong
    if k.b200_graph_launch_count() - g0 > 0:
        launches += 3 * (cg1 - cg0)
    status, obj = r.info.status, r.info.obj_val
    solver.cleanup()

    # ---- end-to-end arm: host arrays -> setup -> solve -> solution on the host, host clock
    def e2e_step():
        t0 = time.perf_counter()
        s = OSQP(prec).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **SETTINGS)
        t1 = time.perf_counter()
        rr = s.solve()
        t2 = time.perf_counter()
        s.cleanup()
        return rr.info.iter, t2 - t0, t1 - t0
    for _ in range(min(args.warmup, 1)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    e_iters, setup_s, e_steps = 0, 0.0, []
    for _ in range(args.steps):
        it, dt, st = e2e_step()
        e_iters += it
        setup_s += st
        e_steps.append(round(1e3 * dt, 1))
    barrier()
    e_total = time.perf_counter() - t0
    fi = F if prec == "f64" else 4
    # uploads of one setup: CSC(A) once (= CSR(A'); CSR(A) is built on the device), the full symmetric
    # P expanded on the host, q / l / u
    h2d = nnzA * (fi + 4) + (n + 1) * 4 + (2 * nnzP + n) * (fi + 4) + (n + 1) * 4 + (n + 2 * m) * fi
    d2h = 2 * (n + m) * fi

    # ---- aggregate over ranks: max time, summed work
    if dist is not None:
        import torch
        t = torch.tensor([ms, e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([iters, e_iters, launches, cg1 - cg0, ns1 - ns0], device="cuda", dtype=torch.float64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ms, e_total = t.tolist()
        iters, e_iters, launches, dcg, dns = w.tolist()
        if sharded:     # every rank ran the SAME iterations of the one shared QP
            iters, e_iters, dcg, dns = iters / world, e_iters / world, dcg / world, dns / world
    else:
        dcg, dns = cg1 - cg0, ns1 - ns0

    if rank == 0:
        peak, peak_src = peaks()
        kcg = max(1, int(round(dcg / max(dns, 1))))
        roof = pcg_roofline(k, pb, kcg) if (prec == "f64" and not sharded) else None
        traffic = None
        tp = ROOT / "profiles" / "pcg_traffic.json"
        if tp.exists() and roof is not None:
            try:
                tj = json.loads(tp.read_text())
                # only comparable for the same kernel (and, for the persistent kernel, the same
                # number of CG iterations per launch)
                if roof["kernel"].startswith(tj.get("kernel", "?")) and (
                        tj.get("cg_iters_per_launch") in (None, roof["cg_iters_per_launch"])):
                    traffic = tj.get("dram_bytes_per_launch")
            except (ValueError, OSError):
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                lib = oracle_library()
                cpb = make_problem(CPU_SAMPLE_SCALE, seed=1)
                it, dt, st, _ = cpu_step(lib, cpb)
                cpu = {"value": it / dt, "unit": "iter/s", "cores": 1, "kind": "reference",
                       "sample": (f"one osqp_setup + osqp_solve of the Lasso generator at scale "
                                  f"{CPU_SAMPLE_SCALE} (n={cpb['P'].shape[0]}, m={cpb['A'].shape[0]}, "
                                  f"nnz(A)={cpb['A'].nnz}) on 1 host core: unmodified reference core + "
                                  "builtin backend + QDLDL restatement; full size does not fit QDLDL"),
                       "solve_only_iters_per_sec": it / st if st else None, "seconds": dt}
            except Exception as exc:   # the oracle is optional for the product arm
                cpu = {"value": None, "unit": "iter/s", "cores": 1, "kind": "reference",
                       "sample": f"unavailable: {exc}"}
        out = {
            "metric": "admm_iters_per_sec", "value": iters / (ms / 1e3), "unit": "iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None, "dtype": prec, "data": "synthetic",
            "config": {"workload": ("lasso_1e5x1e6 (BASELINE configs[1])" if (args.scale == 1.0 and args.workload == "lasso")
                                    else f"{args.workload} generator at scale {args.scale}"),
                       "n": n, "m": m, "nnzA": nnzA, "nnzP": nnzP, "eps": 1e-3,
                       "solver": "indirect: device-resident Jacobi PCG on the reduced KKT system (CUDA-graph WHILE loop of lean sm_100a passes)",
                       "step": "one cold-start osqp_solve to eps 1e-3",
                       "parallelism": ("1 GPU" if world == 1 else
                                       (f"one QP row-sharded over {world} GPUs, column-split layout: 1 NCCL all-reduce of the shared columns + 2 scalar exchanges per CG iteration"
                                        if sharded else f"{world} independent QPs (same generator and seed), one per GPU, no comms")),
                       "l2_policy": "working set (>=460 MB of matrices per CG iteration) exceeds the 126 MB L2",
                       "settings": {kk: vv for kk, vv in SETTINGS.items()}},
            "admm_iters_per_step": iters / args.steps / (1 if sharded else world),
            "cg_iters_per_admm_iter": dcg / max(dns, 1),
            "time_to_solution_ms": ms / args.steps,
            "status": status, "obj_val": obj,
            "e2e": {"value": e_iters / e_total, "unit": "iter/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "time_to_solution_ms": 1e3 * e_total / args.steps,
                    "setup_ms": 1e3 * setup_s / args.steps, "step_ms": e_steps,
                    "step": "osqp_setup from host CSC arrays + osqp_solve + solution to host"},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if roof is not None:
            out["roofline"] = {"bound": "hbm", "achieved": roof["gbs"], "peak": peak, "unit": "GB/s",
                               "frac": roof["gbs"] / peak, "traffic": traffic, "peak_source": peak_src,
                               "kernel": roof["kernel"], "cg_iters_per_launch": roof["cg_iters_per_launch"],
                               "bytes_per_launch": roof["bytes_per_launch"],
                               "ms_per_launch": roof["ms_per_launch"]}
            for extra in ("solve", "pass_A", "phases_us"):
                if extra in roof:
                    out["roofline"][extra] = roof[extra]
        out["cpu_baseline"] = cpu
        args.emit(json.dumps(out))
    if dist is not None:
        dist.barrier()
        if sharded:
            k.b200_dist_finalize()
        dist.destroy_process_group()


class QuietStdout:
    """Everything any library writes to fd 1 (NCCL prints its version there) goes to stderr; the one
    JSON line is written to the real stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        self.out = os.fdopen(self.real, "w")
        return self

    def emit(self, line):
        self.out.write(line + "\n")
        self.out.flush()

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="problem scale (1.0 = BASELINE configs[1])")
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="batch", choices=["batch", "sharded"],
                    help="N > 1: independent QPs per GPU (default, no comms) or ONE row-sharded QP")
    ap.add_argument("--workload", default="lasso",
                    choices=["lasso", "portfolio", "huber", "svm", "random_qp", "mpc"])
    args = ap.parse_args()
    if args.warmup < 1:
        args.warmup = 1
    with QuietStdout() as q:
        args.emit = q.emit
        if args.impl == "reference":
            run_reference(args)
        else:
            run_b200(args)


if __name__ == "__main__":
    main()
