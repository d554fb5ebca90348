"""CPU tests of the host-side mirror (osqp_b200/interface.py) and of the solution assembly of the
row-sharded mode: no GPU, no kernel library."""
import numpy as np
import scipy.sparse as sp

from osqp_b200.dist import assemble_solution, partition_rows
from osqp_b200.interface import _csc_struct, _upper_triangle
from osqp_b200 import _capi


def test_upper_triangle_passes_a_triangle_through_and_cuts_a_full_matrix():
    rng = np.random.default_rng(0)
    M = sp.random(60, 60, density=0.1, format="csc", random_state=1, data_rvs=lambda s: rng.standard_normal(s))
    full = (M + M.T + sp.eye(60)).tocsc()
    tri = sp.triu(full, format="csc")
    out = _upper_triangle(tri)
    assert out is tri                                    # no COO round trip for the common case
    cut = _upper_triangle(full)
    assert (cut != tri).nnz == 0 and (sp.tril(cut, -1)).nnz == 0
    # other formats are converted, empty matrices pass
    assert (_upper_triangle(full.tocsr()) != tri).nnz == 0
    assert _upper_triangle(sp.csc_matrix((5, 5))).nnz == 0


def test_csc_struct_views_the_callers_arrays_and_sorts_indices():
    T = _capi.TYPES_F64
    A = sp.csc_matrix(np.array([[0.0, 2.0], [3.0, 0.0], [4.0, 5.0]]))
    A.indices = A.indices.astype(np.int32)
    A.indptr = A.indptr.astype(np.int32)
    keep = []
    s = _csc_struct(T, A, np.float64, keep)
    assert (s.m, s.n, s.nzmax, s.nz, s.owned) == (3, 2, 4, -1, 0)
    assert [s.p[i] for i in range(3)] == [0, 2, 4] and [s.i[k] for k in range(4)] == [1, 2, 0, 2]
    assert [s.x[k] for k in range(4)] == [3.0, 4.0, 2.0, 5.0]
    # unsorted input comes out sorted (the device transpose relies on ascending rows per column only
    # for bit-identity with the host path, but the reference API requires it: validate_data)
    B = sp.csc_matrix((np.array([1.0, 2.0]), np.array([2, 0]), np.array([0, 2])), shape=(3, 1))
    keep = []
    sB = _csc_struct(T, B, np.float64, keep)
    assert [sB.i[k] for k in range(2)] == [0, 2] and [sB.x[k] for k in range(2)] == [2.0, 1.0]


def test_assemble_solution_row_blocks_layout():
    A = sp.random(40, 7, density=0.4, format="csr", random_state=2)
    b = partition_rows(A, 3)
    x = np.arange(7.0)
    y = np.arange(40.0) * 0.5
    parts = [(x.copy(), y[int(b[r]):int(b[r + 1])]) for r in range(3)]
    xg, yg = assemble_solution(parts, 7, 40, bounds=b)
    assert np.array_equal(xg, x) and np.array_equal(yg, y)
