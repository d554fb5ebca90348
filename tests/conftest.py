import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
GOLDEN = Path(__file__).resolve().parent / "golden"
REF = Path(os.environ.get("OSQP_REFERENCE", "/root/reference"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a box without a GPU: the gpu-marked tests are skipped, not failed.  On a
    box with a GPU nothing is skipped here: a library that does not come up there is a failure."""
    if _gpu_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device on this box (gpu-marked test)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Golden fixture written by tests/golden/make_golden.py from the reference's own
    generate_problem.py scripts; sparse matrices are rebuilt as scipy CSC."""
    raw = np.load(GOLDEN / f"{name}.npz")
    out = {}
    for k in raw.files:
        if k.endswith("__data"):
            b = k[:-6]
            out[b] = sp.csc_matrix((raw[b + "__data"], raw[b + "__indices"], raw[b + "__indptr"]),
                                   shape=tuple(raw[b + "__shape"]))
        elif k.endswith(("__indices", "__indptr", "__shape")):
            continue
        elif k.endswith("__str"):
            out[k[:-5]] = str(raw[k])
        else:
            v = raw[k]
            out[k] = v.item() if v.ndim == 0 else v
    return out


def _ensure_oracle():
    lib = ROOT / "oracle" / "_ref" / "libosqp_builtin.so"
    if not lib.exists() and REF.exists():
        subprocess.run(["make", "-C", str(ROOT / "oracle"), f"REF={REF}"], check=True,
                       stdout=subprocess.DEVNULL)
    return lib


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle: unmodified reference core + builtin backend + local QDLDL restatement."""
    from osqp_b200.interface import LoadedLibrary
    lib = _ensure_oracle()
    if not lib.exists():
        pytest.skip("oracle/_ref/libosqp_builtin.so missing and no reference tree to build it")
    return LoadedLibrary(lib)


@pytest.fixture(scope="session")
def b200_lib():
    """The product library; no fallback -- a missing build is a test failure."""
    from osqp_b200 import load_library
    return load_library("f64")


def _gpu_visible():
    import shutil
    if not shutil.which("nvidia-smi"):
        return False
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
    except (OSError, subprocess.TimeoutExpired):
        return False
    return any(l.startswith("GPU ") for l in out.splitlines())


@pytest.fixture(scope="session")
def kern():
    """The kernel library with a live context on cuda:0.  On a box WITHOUT a GPU the test is skipped
    (`pytest tests` stays usable there); on a box WITH one, a failing b200_init is a hard failure --
    there is no fallback to hide behind."""
    from osqp_b200.devmem import kernels
    k = kernels("f64")
    rc = k.b200_init(0)
    if rc != 0 and not _gpu_visible():
        pytest.skip("no CUDA device on this box")
    assert rc == 0, "b200_init failed although a GPU is visible"
    yield k
    assert k.b200_last_error() == 0, "sticky CUDA error recorded during the session"


# the reference test fixture's common settings (tests/osqp_tester.h:60-82)
FIXTURE_SETTINGS = dict(rho=0.1, alpha=1.6, max_iter=2000, scaling=1, eps_abs=1e-5, eps_rel=1e-5)
TESTS_TOL = 1e-4   # tests/osqp_tester.h:12-16 (double)
