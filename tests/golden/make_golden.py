"""Generate the golden fixtures in this directory FROM THE REFERENCE'S OWN TEST GENERATORS.

Runs /root/reference/tests/<case>/generate_problem.py unmodified (they are plain NumPy/SciPy
scripts seeded with PCG64(1), tests/lin_alg/generate_problem.py:8) with the two C-code emitters
of tests/utils/codegen_utils.py (generate_problem_data :171, generate_data :339) replaced by
capture hooks, and stores every input AND expected answer as <case>.npz.  The reference tree
does not travel to the GPU box, the .npz files do.

    python tests/golden/make_golden.py        # needs /root/reference
"""
import importlib
import os
import sys
import warnings

import numpy as np
import scipy.sparse as sp

REF_TESTS = os.environ.get("OSQP_REF_TESTS", "/root/reference/tests")
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["basic_lp", "basic_qp", "basic_qp2", "lin_alg", "no_active_set", "non_cvx",
         "primal_dual_infeasibility", "primal_infeasibility", "solve_linsys", "unconstrained",
         "update_matrices"]


def flatten(d):
    out = {}
    for k, v in d.items():
        if sp.issparse(v):
            v = sp.csc_matrix(v)
            v.sort_indices()
            out[k + "__data"] = np.asarray(v.data, dtype=np.float64)
            out[k + "__indices"] = np.asarray(v.indices, dtype=np.int32)
            out[k + "__indptr"] = np.asarray(v.indptr, dtype=np.int32)
            out[k + "__shape"] = np.asarray(v.shape, dtype=np.int64)
        elif isinstance(v, str):
            out[k + "__str"] = np.array(v)
        else:
            out[k] = np.asarray(v)
    return out


def main():
    sys.path.insert(0, REF_TESTS)
    warnings.simplefilter("ignore")
    import utils.codegen_utils as cu
    captured = {}

    def cap_problem(P, q, A, l, u, problem_name, sols_data={}):
        d = dict(sols_data)
        d.update(P=P, q=q, A=A, l=l, u=u)
        captured[problem_name] = d

    def cap_data(problem_name, sols_data):
        captured[problem_name] = dict(sols_data)

    cu.generate_problem_data = cap_problem
    cu.generate_data = cap_data
    for case in CASES:
        importlib.import_module(case + ".generate_problem")
    for name, d in captured.items():
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **flatten(d))
        print(f"{name}: {len(d)} entries -> {path}")


if __name__ == "__main__":
    main()
