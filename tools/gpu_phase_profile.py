"""Development probe: where does one ADMM iteration of the headline workload spend its time?

For each PCG driver (persistent kernel / graph) in a fresh process: set the Lasso workload up with
the bench settings, solve, report time per solve, iteration counts, the launch trace of the last
solve (B200_TRACE_FILE, see csrc/context.cu) and -- graph driver only -- the per-kernel phase
profile (b200_pcg_profile_last).
"""
import argparse
import collections
import csv
import ctypes as C
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, ".")

SETTINGS = dict(eps_abs=1e-3, eps_rel=1e-3, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
                polishing=0, verbose=0, warm_starting=0)


def child(args):
    import numpy as np
    from osqp_b200 import OSQP, problems
    from osqp_b200.devmem import kernels
    k = kernels("f64")
    assert k.b200_init(0) == 0
    if args.family == "lasso":
        pb = problems.lasso(int(1e5 * args.scale), int(1e6 * args.scale))
    elif args.family == "svm":
        pb = problems.svm(int(1e4), int(1e7 * args.scale))
    elif args.family == "huber":
        pb = problems.huber(int(1e4), int(1e7 * args.scale))
    else:
        pb = problems.random_qp()
    s = OSQP("f64").setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **SETTINGS)
    e0, e1 = k.b200_event_create(), k.b200_event_create()
    out = {"driver": os.environ.get("B200_PCG_DRIVER", "persistent"), "solves": []}
    for rep in range(args.reps):
        cg0, ns0 = s.cg_stats()
        l0 = k.b200_launch_count()
        k.b200_event_record(e0)
        r = s.solve()
        k.b200_event_record(e1)
        ms = k.b200_event_elapsed_ms(e0, e1)
        cg1, ns1 = s.cg_stats()
        out["solves"].append(dict(ms=round(ms, 3), iters=r.info.iter, status=r.info.status, obj=r.info.obj_val,
                                  cg=cg1 - cg0, lin_solves=ns1 - ns0, launches=k.b200_launch_count() - l0))
    if os.environ.get("B200_PCG_DRIVER") == "graph":
        k.b200_pcg_profile_last.argtypes = [C.c_int, C.POINTER(C.c_double), C.c_int]
        k.b200_pcg_profile_last.restype = C.c_int
        buf = (C.c_double * 14)()
        rc = k.b200_pcg_profile_last(20, buf, 14)
        names = ["lean_A(L1)", "lean_K2(L2,3dots)", "update_fused(L3+L4)", "seq A+K2+update", "generic K<0>(P2)",
                 "lean K2 P2", "p1_carried", "epilogue", "(retired)", "(retired) ", "nop launch",
                 "lean A exact", "lean At plain", "lean K2 plain"]
        out["phase_us"] = {nm: round(buf[i], 2) for i, nm in enumerate(names)} if rc == 0 else f"rc={rc}"
    s.cleanup()     # last solver alive: b200_shutdown dumps the launch trace
    k.b200_shutdown()
    print("RESULT " + json.dumps(out), flush=True)


def summarise_trace(path, last_launches):
    rows = []
    for ln in open(path).read().splitlines()[1:]:
        idx, rest = ln.split(",", 1)
        tag, dev, host = rest.rsplit(",", 2)        # tags may contain commas (template arguments)
        rows.append((tag[:48], float(dev), float(host)))
    rows = rows[-last_launches:]
    agg = collections.OrderedDict()
    for tag, dev, host in rows:
        a = agg.setdefault(tag, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += dev
        a[2] += host
    tot = sum(a[1] for a in agg.values())
    print(f"  launch trace of the last solve ({len(rows)} launches, {tot/1e3:.2f} ms device time):")
    for tag, (cnt, dev, host) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"    {tag:48s} n={cnt:5d}  dev {dev/1e3:8.2f} ms ({100*dev/tot:5.1f} %)  {dev/cnt:8.1f} us each   host {host/cnt:7.1f} us each")


def summarise_marked(path):
    """Rows between the last solve-begin / solve-end markers of a trace."""
    lines = open(path).read().splitlines()[1:]
    tags = [ln.split(",", 1)[1].rsplit(",", 2)[0] for ln in lines]
    b = max(i for i, t in enumerate(tags) if t == "solve-begin")
    e = max(i for i, t in enumerate(tags) if t == "solve-end")
    tmp = path + ".last"
    open(tmp, "w").write("hdr\n" + "\n".join(lines[b + 1:e]) + "\n")
    summarise_trace(tmp, e - b - 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--family", default="lasso")
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--drivers", default="graph,persistent")
    ap.add_argument("--summarise", default=None, help="trace CSV with solve-begin/solve-end markers")
    args = ap.parse_args()
    if args.summarise:
        return summarise_marked(args.summarise)
    if args.child:
        return child(args)
    os.makedirs("gpurun_out", exist_ok=True)
    for drv in args.drivers.split(","):
        env = dict(os.environ)
        if drv == "graph":
            env["B200_PCG_DRIVER"] = "graph"
        else:
            env.pop("B200_PCG_DRIVER", None)
        trace = f"gpurun_out/trace_{args.family}_{drv}.csv"
        env["B200_TRACE_FILE"] = trace
        t0 = time.time()
        p = subprocess.run([sys.executable, __file__, "--child", "--family", args.family, "--scale", str(args.scale),
                            "--reps", str(args.reps)], env=env, capture_output=True, text=True)
        res = [ln for ln in p.stdout.splitlines() if ln.startswith("RESULT ")]
        print(f"== driver {drv} ({time.time()-t0:.0f} s, rc {p.returncode})", flush=True)
        if not res:
            print(p.stdout[-2000:], p.stderr[-3000:])
            continue
        out = json.loads(res[0][7:])
        for sv in out["solves"]:
            print("  ", sv)
        if "phase_us" in out:
            print("  phase_us:", json.dumps(out["phase_us"], indent=4))
        if os.path.exists(trace):
            summarise_trace(trace, out["solves"][-1]["launches"])


if __name__ == "__main__":
    main()
