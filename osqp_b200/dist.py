"""Row-sharded multi-GPU mode: host-side partitioning and process-group plumbing.

One process per GPU (torchrun).  Every rank runs the SAME unmodified OSQP core on its row block
(A_r, l_r, u_r) with P and q replicated; the backend combines what has to be combined
(algebra/b200: reductions over m-vectors, A'y, column norms; osqp_b200/csrc/pcg_graph.cu: one
all-reduce of the length-n partial K p per CG iteration).  torch.distributed is used only to hand
the 128-byte NCCL unique id from rank 0 to the other ranks (any backend, gloo included).
"""
import ctypes as C

import numpy as np
import scipy.sparse as sp

from . import OSQP
from .devmem import kernels


def partition_rows(A, world, n=None):
    """Contiguous row blocks balanced by nonzeros (+1 per row for the vector work).
    Returns the world+1 row boundaries.  No block may have exactly n rows (the backend tells
    row-sharded vectors from replicated ones by their length), and none may be empty."""
    A = sp.csr_matrix(A)
    m = A.shape[0]
    n = A.shape[1] if n is None else n
    if world <= 1:
        return np.array([0, m], dtype=np.int64)
    if m < world:
        raise ValueError("fewer rows than ranks")
    w = np.diff(A.indptr).astype(np.int64) + 1
    cum = np.concatenate([[0], np.cumsum(w)])
    bounds = [0]
    for r in range(1, world):
        b = int(np.searchsorted(cum, cum[-1] * r / world))
        b = min(max(b, bounds[-1] + 1), m - (world - r))
        bounds.append(b)
    bounds.append(m)
    bounds = np.array(bounds, dtype=np.int64)
    for r in range(world):                       # nudge a boundary if a block has exactly n rows
        if bounds[r + 1] - bounds[r] == n:
            if r + 1 < world and bounds[r + 2] - bounds[r + 1] > 1:
                bounds[r + 1] += 1
            elif r > 0 and bounds[r] - bounds[r - 1] > 1:
                bounds[r] -= 1
            else:
                raise ValueError("cannot avoid a block with exactly n rows")
    assert (np.diff(bounds) > 0).all() and not (np.diff(bounds) == n).any()
    return bounds


def shard_problem(pb, rank, world):
    """This rank's view of the QP: P, q replicated; rows [b_r, b_{r+1}) of A, l, u."""
    A = sp.csr_matrix(pb["A"])
    bounds = partition_rows(A, world, n=A.shape[1])
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    return dict(P=pb["P"], q=pb["q"], A=A[r0:r1].tocsc(), l=np.asarray(pb["l"])[r0:r1],
                u=np.asarray(pb["u"])[r0:r1], rows=(r0, r1), bounds=bounds)


def exchange_unique_id(make_id, dist):
    """rank 0 creates the NCCL unique id, everybody receives it (works on any backend)."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8)
    if dist.get_rank() == 0:
        buf = torch.from_numpy(np.frombuffer(make_id(), dtype=np.uint8).copy())
    if dist.get_backend() == "nccl":
        buf = buf.cuda()
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def init_sharded(dist, local_rank, precision="f64"):
    """b200_init on cuda:local_rank + NCCL communicator over all ranks of `dist`."""
    k = kernels(precision)
    if k.b200_init(local_rank) != 0:
        raise RuntimeError("no usable GPU: the B200 backend has no CPU fallback")
    k.b200_dist_unique_id.argtypes = [C.c_char_p]
    k.b200_dist_unique_id.restype = C.c_int
    k.b200_dist_init.argtypes = [C.c_int, C.c_int, C.c_char_p]
    k.b200_dist_init.restype = C.c_int

    def make_id():
        raw = C.create_string_buffer(128)
        if k.b200_dist_unique_id(raw) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
        return raw.raw
    uid = exchange_unique_id(make_id, dist)
    if k.b200_dist_init(dist.get_rank(), dist.get_world_size(), uid) != 0:
        raise RuntimeError("ncclCommInitRank failed")
    return k


class ShardedOSQP(OSQP):
    """OSQP solver object over a row-sharded problem.  `setup` takes the FULL problem on every
    rank and keeps this rank's row block; `solve` returns x (replicated) and the local slice of y."""

    def __init__(self, rank, world, precision="f64"):
        super().__init__(precision)
        self.rank, self.world = rank, world
        self._lib.osqp_b200_dist_configure.argtypes = [C.c_int, C.c_int]
        self._lib.osqp_b200_dist_configure.restype = C.c_int

    def setup(self, P, q, A, l, u, **settings):
        sh = shard_problem(dict(P=P, q=q, A=A, l=l, u=u), self.rank, self.world)
        self.rows, self.bounds = sh["rows"], sh["bounds"]
        n = sp.csc_matrix(P).shape[0]
        if self._lib.osqp_b200_dist_configure(n, sh["A"].shape[0]) != 0:
            raise ValueError("a row block may not have exactly n rows")
        return super().setup(sh["P"], sh["q"], sh["A"], sh["l"], sh["u"], **settings)

    def cleanup(self):
        super().cleanup()
        try:
            self._lib.osqp_b200_dist_configure(0, -1)
        except Exception:
            pass
