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


def partition_rows(A, world, n=None, row_nnz=None):
    """Contiguous row blocks balanced by nonzeros (+1 per row for the vector work).
    Returns the world+1 row boundaries.  No block may have exactly n rows (the backend tells
    row-sharded vectors from replicated ones by their length), and none may be empty.
    `row_nnz` (entries per row) spares the CSR conversion when the caller already has it."""
    m = A.shape[0]
    n = A.shape[1] if n is None else n
    if world <= 1:
        return np.array([0, m], dtype=np.int64)
    if m < world:
        raise ValueError("fewer rows than ranks")
    if row_nnz is None:
        row_nnz = np.diff(sp.csr_matrix(A).indptr)
    w = np.asarray(row_nnz).astype(np.int64) + 1
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


# --------------------------------------------------------------------------------------------
# Shared / local column split (SURVEY.md section 8e "recommended refinement")
#
# With plain row blocks every n-vector is replicated, so the BLAS-1 work and the all-reduce volume
# do not shrink with the number of ranks -- fatal for epigraph-style problems (Huber, SVM, Lasso)
# whose slack variables outnumber the features 1000:1.  Here rows that are linked through
# low-degree columns (a slack variable and the 2-3 rows it appears in) are first clustered with a
# union-find so that they land on the same rank; a column is then LOCAL to a rank if all its
# nonzeros (and its couplings in P) live there, and SHARED otherwise.  Each rank solves over
# [shared columns ; its local columns]: only the shared slice is ever exchanged.

def plan_column_split(P, A, world, max_link_degree=3):
    """Returns a dict with, per rank r: rows[r] (global row ids), cols[r] = shared ++ local_r (global
    column ids), and `shared` (global ids of the shared columns, identical leading part of cols[r])."""
    A = sp.csc_matrix(A)
    m, n = A.shape
    Ptri = sp.triu(sp.csc_matrix(P), k=1).tocoo()
    deg = np.diff(A.indptr)
    # 1. cluster rows through low-degree columns (connected components of the row graph whose
    #    edges join the first row of every linking column to its other rows)
    from scipy.sparse.csgraph import connected_components
    link = np.nonzero((deg >= 2) & (deg <= max_link_degree))[0]
    if link.size:
        starts = A.indptr[link]
        src, dst = [], []
        for t in range(1, max_link_degree):
            has = deg[link] > t
            src.append(A.indices[starts[has]])
            dst.append(A.indices[starts[has] + t])
        src, dst = np.concatenate(src), np.concatenate(dst)
        G = sp.coo_matrix((np.ones(src.size, dtype=np.int8), (src, dst)), shape=(m, m))
        _, comp = connected_components(G, directed=False)
    else:
        comp = np.arange(m)
    # label every cluster by its first row so that clusters keep the original row order
    o = np.argsort(comp, kind="stable")                     # stable: the first row of a cluster comes first
    starts_c = np.concatenate([[0], np.nonzero(np.diff(comp[o]))[0] + 1])
    first = np.empty(comp.max() + 1, dtype=np.int64)
    first[comp[o[starts_c]]] = o[starts_c]
    root = first[comp]
    # 2. clusters -> ranks, greedy in order of first row, balanced by nonzeros (+1 per row)
    # entries per row straight from the CSC row indices: the planner never needs a CSR copy
    row_nnz = np.bincount(A.indices, minlength=m)
    w_row = row_nnz.astype(np.int64) + 1
    order = np.argsort(root, kind="stable")                 # rows grouped by cluster, clusters by first row
    w_sorted = w_row[order]
    cum = np.cumsum(w_sorted)
    target = cum[-1] / world
    rank_sorted = np.minimum((cum - w_sorted) // max(target, 1), world - 1).astype(np.int64)
    # a cluster must not straddle two ranks: every row takes the rank of its cluster's first row
    first_of_cluster = np.concatenate([[True], root[order][1:] != root[order][:-1]])
    start_idx = np.maximum.accumulate(np.where(first_of_cluster, np.arange(len(order)), 0))
    cl_rank = rank_sorted[start_idx]
    rank_of_row = np.empty(m, dtype=np.int64)
    rank_of_row[order] = cl_rank
    # 3. classify columns
    owner = np.full(n, -1, dtype=np.int64)
    r_entry = rank_of_row[A.indices]
    lo = np.full(n, world, dtype=np.int64)
    hi = np.full(n, -1, dtype=np.int64)
    ne = np.nonzero(deg > 0)[0]                             # CSC: the entries of a column are contiguous
    if ne.size:
        lo[ne] = np.minimum.reduceat(r_entry, A.indptr[ne])
        hi[ne] = np.maximum.reduceat(r_entry, A.indptr[ne])
    single = (deg > 0) & (lo == hi)
    owner[single] = lo[single]
    # columns coupled through P must live together: shared wins, propagate until stable
    if Ptri.nnz:
        changed = True
        while changed:
            bad = owner[Ptri.row] != owner[Ptri.col]
            touch = np.unique(np.concatenate([Ptri.row[bad], Ptri.col[bad]]))
            touch = touch[owner[touch] >= 0]
            changed = touch.size > 0
            owner[touch] = -1
    shared = np.nonzero(owner < 0)[0]
    # quality guard: if the cut is lopsided (an empty or 2x overloaded rank) or barely anything is
    # local, fall back to plain contiguous row blocks with every column shared
    cnt = np.bincount(rank_of_row, weights=w_row, minlength=world)
    if cnt.min() == 0 or cnt.max() > 2.0 * cnt.mean() or shared.size > 0.5 * n:
        bounds = partition_rows(A, world, n=-1, row_nnz=row_nnz)
        rank_of_row = np.repeat(np.arange(world), np.diff(bounds))
        owner[:] = -1
        shared = np.arange(n)
    rows, cols = [], []
    row_local = np.empty(m, dtype=np.int64)          # position of a row inside its rank's block
    for r in range(world):
        rows.append(np.nonzero(rank_of_row == r)[0])
        row_local[rows[r]] = np.arange(rows[r].size)
        cols.append(np.concatenate([shared, np.nonzero(owner == r)[0]]))
    return dict(shared=shared, rows=rows, cols=cols, n=n, m=m, world=world, rank_of_row=rank_of_row,
                row_local=row_local)


def shard_problem_split(pb, rank, plan, Acsr=None):
    """Rank `rank`'s QP under a column-split plan.  A row with no entries and infinite bounds is
    appended when the local problem would otherwise have as many rows as columns (the backend
    tells row vectors from column vectors by their length)."""
    P = sp.csc_matrix(pb["P"])
    R, Cc = plan["rows"][rank], plan["cols"][rank]
    if Acsr is not None:
        A_r = Acsr[R][:, Cc].tocsc()
    else:
        A_r = _slice_csc(pb["A"], R.size, Cc, plan["rank_of_row"], plan["row_local"], rank)
    P_r = sp.triu(P[Cc][:, Cc], format="csc")
    l, u = np.asarray(pb["l"], dtype=float)[R], np.asarray(pb["u"], dtype=float)[R]
    padded = 0
    if A_r.shape[0] == A_r.shape[1]:
        A_r = sp.vstack([A_r, sp.csc_matrix((1, A_r.shape[1]))], format="csc")
        l, u = np.append(l, -np.inf), np.append(u, np.inf)
        padded = 1
    return dict(P=P_r, q=np.asarray(pb["q"], dtype=float)[Cc], A=A_r, l=l, u=u, padded=padded,
                n_shared=int(plan["shared"].size))


def _slice_csc(A, n_rows_local, cols, rank_of_row, row_local, rank):
    """A[rows of `rank`][:, cols] as CSC without any format conversion: gather the selected columns'
    segments, keep the entries whose row lives on `rank`, renumber the rows.  Row order inside a
    column is preserved (row_local is increasing in the global row id), so sorted input gives sorted
    output -- identical to scipy's A.tocsr()[rows][:, cols].tocsc()."""
    A = A if sp.isspmatrix_csc(A) else sp.csc_matrix(A)
    A.sort_indices()
    lengths = (A.indptr[cols + 1] - A.indptr[cols]).astype(np.int64)
    total = int(lengths.sum())
    starts = A.indptr[cols].astype(np.int64)
    # positions of the selected entries in A.indices / A.data, column after column
    offs = np.concatenate([[0], np.cumsum(lengths)])[:-1]
    idx = np.repeat(starts - offs, lengths) + np.arange(total, dtype=np.int64)
    rows_g = A.indices[idx]
    keep = rank_of_row[rows_g] == rank
    col_id = np.repeat(np.arange(cols.size, dtype=np.int64), lengths)[keep]
    indptr = np.concatenate([[0], np.cumsum(np.bincount(col_id, minlength=cols.size))])
    out = sp.csc_matrix((A.data[idx][keep], row_local[rows_g[keep]].astype(np.int32), indptr.astype(np.int32)),
                        shape=(n_rows_local, cols.size))
    out.has_sorted_indices = True
    return out


def assemble_solution(parts, n_global, m_global, plan=None, bounds=None):
    """Global (x, y) from the per-rank (x_r, y_r) pairs.  Column-split layout (`plan`): x_r is
    [shared slice ; owned slice] in the order of plan["cols"][r] (the shared slice is taken from rank
    0), y_r the rows plan["rows"][r] (a padded free row, if any, is dropped).  Plain row blocks
    (`bounds`): x replicated, y_r the contiguous block of rank r."""
    x, y = np.zeros(n_global), np.zeros(m_global)
    if plan is not None:
        ns = plan["shared"].size
        for rk, (xr, yr) in enumerate(parts):
            rows, cols = plan["rows"][rk], plan["cols"][rk]
            if rk == 0:
                x[cols] = xr
            else:
                x[cols[ns:]] = xr[ns:]
            y[rows] = yr[:len(rows)]
    else:
        x[:] = parts[0][0]
        for rk, (_, yr) in enumerate(parts):
            y[int(bounds[rk]):int(bounds[rk + 1])] = yr
    return x, y


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
    enable_peer_exchange(k, dist)
    return k


def enable_peer_exchange(k, dist):
    """Peer-memory exchange for the CG loop (csrc/dist.cu): every rank exports the CUDA IPC handle of its
    exchange buffer, the 64-byte handles are all-gathered in rank order and imported.  Returns True when
    every rank succeeded; otherwise the sharded solve keeps its NCCL path (also with B200_DIST_NO_P2P=1)."""
    import os
    world = dist.get_world_size()
    if world < 2 or world > 8 or os.environ.get("B200_DIST_NO_P2P"):
        return False
    k.b200_dist_p2p_export.argtypes = [C.c_char_p]
    k.b200_dist_p2p_export.restype = C.c_int
    k.b200_dist_p2p_import.argtypes = [C.c_char_p, C.c_int]
    k.b200_dist_p2p_import.restype = C.c_int
    raw = C.create_string_buffer(64)
    rc = k.b200_dist_p2p_export(raw)
    got = [None] * world
    dist.all_gather_object(got, (int(rc), raw.raw))
    if any(r != 0 for r, _ in got):
        return False
    rc = k.b200_dist_p2p_import(b"".join(h for _, h in got), world)
    oks = [None] * world
    dist.all_gather_object(oks, int(rc))
    return all(o == 0 for o in oks)


class ShardedOSQP(OSQP):
    """OSQP solver object over ONE QP whose rows of A are split over the ranks.

    layout="split" (default): column-split layout -- rows linked through low-degree columns are kept
    together, every rank solves over [shared columns ; the columns only it touches], and only the
    shared slice is ever exchanged (plan_column_split).  layout="rows": plain contiguous row blocks
    with every n-vector replicated.  `setup` takes the FULL problem on every rank; `solve` returns
    this rank's slice of the solution, `gather` assembles the global (x, y)."""

    def __init__(self, rank, world, precision="f64", layout="split"):
        super().__init__(precision)
        self.rank, self.world, self.layout = rank, world, layout
        self.plan = None
        self._lib.osqp_b200_dist_configure.argtypes = [C.c_int, C.c_int]
        self._lib.osqp_b200_dist_configure.restype = C.c_int
        self._lib.osqp_b200_dist_configure_split.argtypes = [C.c_int] * 4
        self._lib.osqp_b200_dist_configure_split.restype = C.c_int

    @staticmethod
    def _check_settings(settings):
        """Every rank must take the same host-side decisions: options whose control flow depends on a
        rank-local clock or that run kernels over vectors of other lengths are refused."""
        if settings.get("polishing", 0):
            raise ValueError("polishing is not supported in the row-sharded mode")
        if settings.get("time_limit", 0):
            raise ValueError("time_limit makes ranks stop at different iterations: not supported when sharded")
        if settings.get("adaptive_rho", 1) not in (0, 1):
            raise ValueError("row-sharded mode needs adaptive_rho in {0 (off), 1 (iteration based)}")

    def setup_local(self, shard, n_global, m_global, **settings):
        """Set up from data that is ALREADY sharded: `shard` is this rank's QP in the column-split layout
        (dict P, q, A, l, u, n_shared as produced by shard_problem_split or problems.*_shard): columns
        [the n_shared columns touched by several ranks ; the columns only this rank touches], its rows
        of A.  No rank ever sees the whole problem."""
        self._check_settings(settings)
        self.n_global, self.m_global = int(n_global), int(m_global)
        A = sp.csc_matrix(shard["A"])
        l, u = np.asarray(shard["l"], dtype=float), np.asarray(shard["u"], dtype=float)
        self.padded = 0
        if A.shape[0] == A.shape[1]:        # the backend tells row vectors from column vectors by length
            A = sp.vstack([A, sp.csc_matrix((1, A.shape[1]))], format="csc")
            l, u = np.append(l, -np.inf), np.append(u, np.inf)
            self.padded = 1
        self.n_shared = int(shard["n_shared"])
        rc = self._lib.osqp_b200_dist_configure_split(A.shape[1], A.shape[0], self.n_shared, self.n_global)
        if rc != 0:
            raise ValueError("invalid column-split layout for this rank")
        return OSQP.setup(self, shard["P"], shard["q"], A, l, u, **settings)

    def setup(self, P, q, A, l, u, **settings):
        self._check_settings(settings)
        pb = dict(P=P, q=q, A=A, l=l, u=u)
        self.n_global, self.m_global = sp.csc_matrix(P).shape[0], sp.csc_matrix(A).shape[0]
        if self.layout == "split":
            # planner and slicer work on the CSC arrays directly: no CSC -> CSR conversion
            self.plan = plan_column_split(P, A, self.world)
            sh = shard_problem_split(pb, self.rank, self.plan)
            self.padded = sh["padded"]
            self.n_shared = sh["n_shared"]
            rc = self._lib.osqp_b200_dist_configure_split(sh["A"].shape[1], sh["A"].shape[0], self.n_shared,
                                                          self.n_global)
            if rc != 0:
                raise ValueError("invalid column-split layout for this rank")
        else:
            sh = shard_problem(pb, self.rank, self.world)
            self.rows, self.bounds = sh["rows"], sh["bounds"]
            if self._lib.osqp_b200_dist_configure(self.n_global, sh["A"].shape[0]) != 0:
                raise ValueError("a row block may not have exactly n rows")
        return super().setup(sh["P"], sh["q"], sh["A"], sh["l"], sh["u"], **settings)

    def gather(self, r, dist):
        """Global (x, y) from the per-rank results `r` (collective; every rank gets the same arrays)."""
        parts = [None] * self.world
        dist.all_gather_object(parts, (np.asarray(r.x), np.asarray(r.y)))
        return assemble_solution(parts, self.n_global, self.m_global, plan=self.plan,
                                 bounds=getattr(self, "bounds", None))

    def cleanup(self):
        super().cleanup()
        try:
            self._lib.osqp_b200_dist_configure(0, -1)
        except Exception:
            pass
