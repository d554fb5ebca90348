"""Independent QPs solved concurrently from several host threads (BASELINE configs[4]: batches of MPC
QPs, no communication).  Every thread owns a library context (stream, workspace, result mailbox), so
the solves overlap on the device; results must be bit-identical to the same solves run one after the
other on a single thread."""
import threading

import numpy as np
import pytest

from osqp_b200 import OSQP, problems

pytestmark = pytest.mark.gpu

SETTINGS = dict(eps_abs=1e-4, eps_rel=1e-4, rho_is_vec=0, adaptive_rho_tolerance=2.0, check_termination=5,
                verbose=0, warm_starting=0)


def _solve(seed, driver=None):
    pb = problems.mpc(N=12, seed=seed)
    s = OSQP("f64").setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **SETTINGS)
    r = s.solve()
    s.cleanup()
    return r


def test_concurrent_threads_match_sequential_solves():
    seeds = list(range(1, 13))
    ref = {sd: _solve(sd) for sd in seeds}
    out, errs = {}, []

    def worker(my):
        try:
            for sd in my:
                out[sd] = _solve(sd)
        except Exception as exc:      # surfaced below: a thread must not die silently
            errs.append(exc)
    threads = [threading.Thread(target=worker, args=(seeds[i::4],)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errs, errs
    for sd in seeds:
        assert out[sd].info.status == ref[sd].info.status == "solved"
        assert out[sd].info.iter == ref[sd].info.iter
        assert np.array_equal(out[sd].x, ref[sd].x) and np.array_equal(out[sd].y, ref[sd].y)


def test_large_and_small_solver_in_parallel_threads():
    """A graph-driver solve (its argument block lives in per-context __constant__ memory) next to
    persistent-kernel solves on other threads."""
    pb = problems.lasso(4000, 100000, density=0.005, seed=3)     # > 2e6 stored entries -> graph driver
    kw = dict(SETTINGS, eps_abs=1e-3, eps_rel=1e-3)
    ref = OSQP("f64").setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
    r0 = ref.solve()
    ref.cleanup()
    res, errs = {}, []

    def big():
        try:
            s = OSQP("f64").setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **kw)
            first = s.solve()          # compare FIRST solves: rho adapted by a solve persists into the next
            s.solve()
            res["big"] = first
            s.cleanup()
        except Exception as exc:
            errs.append(exc)

    def small():
        try:
            res["small"] = [_solve(sd) for sd in (1, 2, 3, 4, 5, 6)]
        except Exception as exc:
            errs.append(exc)
    ts = [threading.Thread(target=big), threading.Thread(target=big), threading.Thread(target=small)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    assert res["big"].info.status == r0.info.status and res["big"].info.iter == r0.info.iter
    assert np.array_equal(res["big"].x, r0.x)
    assert all(r.info.status == "solved" for r in res["small"])
