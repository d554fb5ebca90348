"""CPU checks of the BASELINE-size oracle fixtures (tests/golden/baseline_*.npz): every fixture the
GPU parity tests use exists, matches its seeded generator, and -- for the cases the oracle solves in
seconds -- is reproduced bit-for-bit in status / iteration count and to 1e-9 in the solution by
running the oracle again here."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
import make_baseline_golden as G   # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("case", G.CASES)
def test_fixture_matches_generator(case):
    fx = dict(np.load(GOLDEN / f"baseline_{case}.npz"))
    if case in ("random_qp_full", "mpc_N12", "lasso_s002"):      # the others take seconds each to generate
        assert np.allclose(G.fingerprint(G.build(case)), fx["fingerprint"], rtol=1e-12, atol=0)
    for tag in ("bench",) + (("tight",) if case != "portfolio_mid" else ("nogap",)):
        assert f"{tag}_x" in fx and f"{tag}_x_stride" in fx and fx[f"{tag}_x"].size <= G.MAX_STORED
        assert np.isfinite(fx[f"{tag}_obj"])


@pytest.mark.parametrize("case", ["mpc_N12", "lasso_s002"])
def test_oracle_reproduces_fixture(oracle_lib, case):
    from osqp_b200.interface import OSQP
    fx = dict(np.load(GOLDEN / f"baseline_{case}.npz"))
    pb = G.build(case)
    for tag, st in (("bench", G.BENCH), ("tight", G.TIGHT)):
        r = OSQP(oracle_lib).setup(pb["P"], pb["q"], pb["A"], pb["l"], pb["u"], **st).solve()
        assert r.info.status_val == int(fx[f"{tag}_status"]) and r.info.iter == int(fx[f"{tag}_iter"])
        assert abs(r.info.obj_val - float(fx[f"{tag}_obj"])) <= 1e-9 * max(1.0, abs(r.info.obj_val))
        assert np.abs(np.asarray(r.x)[::int(fx[f"{tag}_x_stride"])] - fx[f"{tag}_x"]).max() <= 1e-9
