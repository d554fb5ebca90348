"""Synthetic QP generators for the BASELINE.json workloads.

Each follows the formulation of the corresponding reference documentation example
(/root/reference/docs/examples/{lasso,portfolio,huber,svm,mpc}.rst) scaled to the sizes
BASELINE.json names; `random_qp` restates the external osqp_benchmarks "Random QP" family
(not in the reference tree).  All return a dict with
P (upper triangle is what the solver uses), q, A, l, u in scipy CSC / numpy float64.
"""
import numpy as np
import scipy.sparse as sp


def sprand(m, n, density, rng, randn=False):
    """Sparse random m x n matrix with ~density*m*n entries, O(nnz) time and memory
    (scipy.sparse.random samples without replacement from m*n, too slow at 1e11 cells)."""
    nnz = int(round(density * m * n))
    rows = rng.integers(0, m, size=nnz, dtype=np.int64)
    cols = rng.integers(0, n, size=nnz, dtype=np.int64)
    key = np.unique(rows * n + cols)
    rows, cols = key // n, key % n
    vals = rng.standard_normal(key.size) if randn else rng.random(key.size)
    return sp.csc_matrix((vals, (rows.astype(np.int32), cols.astype(np.int32))), shape=(m, n))


def random_qp(n=10_000, m=20_000, nnz_target=200_000, seed=1):
    """P = M M' + 1e-2 I, A = sprandn(m, n), u = A v + U(0,1), l = -inf."""
    rng = np.random.default_rng(seed)
    # split the budget: ~75 % in A, rest in triu(P) (M M' roughly squares the row density)
    dA = 0.75 * nnz_target / (m * n)
    A = sprand(m, n, dA, rng, randn=True)
    dM = max(1.0, np.sqrt(0.25 * nnz_target * 2 / n)) / n * 0.7
    M = sprand(n, n, dM, rng, randn=True)
    P = (M @ M.T + 1e-2 * sp.eye(n)).tocsc()
    q = rng.standard_normal(n)
    v = rng.standard_normal(n)
    u = A @ v + rng.random(m)
    l = -np.inf * np.ones(m)
    return dict(P=P, q=q, A=A.tocsc(), l=l, u=u, name=f"random_qp_n{n}_m{m}")


def lasso(n_features=100_000, n_samples=1_000_000, density=1e-4, gamma=None, seed=1):
    """docs/examples/lasso.rst:41-61; variables (x, y, t)."""
    rng = np.random.default_rng(seed)
    n, m = n_features, n_samples
    Ad = sprand(m, n, density, rng)
    x_true = (rng.random(n) > 0.8).astype(float) * rng.standard_normal(n) / np.sqrt(n)
    b = Ad @ x_true + 0.5 * rng.standard_normal(m)
    if gamma is None:
        gamma = np.linspace(1, 10, 11)[5]
    In, Im = sp.eye(n, format="csc"), sp.eye(m, format="csc")
    On = sp.csc_matrix((n, n))
    P = sp.block_diag([On, Im, On], format="csc")
    q = np.hstack([np.zeros(n + m), gamma * np.ones(n)])
    A = sp.bmat([[Ad, -Im, None], [In, None, -In], [In, None, In]], format="csc")
    l = np.hstack([b, -np.inf * np.ones(n), np.zeros(n)])
    u = np.hstack([b, np.zeros(n), np.inf * np.ones(n)])
    return dict(P=P, q=q, A=A, l=l, u=u, name=f"lasso_n{n}_m{m}")


def portfolio(n_assets=1_000_000, k_factors=10_000, density=1e-2, gamma=1.0, seed=1):
    """docs/examples/portfolio.rst:48-63; variables (x, y)."""
    rng = np.random.default_rng(seed)
    n, k = n_assets, k_factors
    F = sprand(n, k, density, rng)
    D = sp.diags(rng.random(n) * np.sqrt(k), format="csc")
    mu = rng.standard_normal(n)
    P = sp.block_diag([D, sp.eye(k)], format="csc")
    q = np.hstack([-mu / (2 * gamma), np.zeros(k)])
    A = sp.bmat([[F.T, -sp.eye(k)],
                 [sp.csc_matrix(np.ones((1, n))), None],
                 [sp.eye(n), None]], format="csc")
    l = np.hstack([np.zeros(k), 1.0, np.zeros(n)])
    u = np.hstack([np.zeros(k), 1.0, np.ones(n)])
    return dict(P=P, q=q, A=A, l=l, u=u, name=f"portfolio_n{n}_k{k}")


def huber(n_features=10_000, n_samples=10_000_000, density=1e-3, seed=1):
    """docs/examples/huber.rst:44-63; variables (x, u, r, s)."""
    rng = np.random.default_rng(seed)
    n, m = n_features, n_samples
    Ad = sprand(m, n, density, rng)
    x_true = rng.standard_normal(n) / np.sqrt(n)
    ind95 = (rng.random(m) < 0.95).astype(float)
    b = Ad @ x_true + 0.5 * rng.standard_normal(m) * ind95 + 10.0 * rng.random(m) * (1.0 - ind95)
    Im = sp.eye(m, format="csc")
    P = sp.block_diag([sp.csc_matrix((n, n)), 2 * Im, sp.csc_matrix((2 * m, 2 * m))], format="csc")
    q = np.append(np.zeros(m + n), 2 * np.ones(2 * m))
    A = sp.bmat([[Ad, -Im, -Im, Im], [None, None, Im, None], [None, None, None, Im]], format="csc")
    l = np.hstack([b, np.zeros(2 * m)])
    u = np.hstack([b, np.inf * np.ones(2 * m)])
    return dict(P=P, q=q, A=A, l=l, u=u, name=f"huber_n{n}_m{m}")


def svm(n_features=10_000, n_samples=10_000_000, density=1e-3, gamma=1.0, seed=1):
    """docs/examples/svm.rst:37-57; variables (x, t)."""
    rng = np.random.default_rng(seed)
    n, m = n_features, n_samples
    N = m // 2
    m = 2 * N
    b = np.hstack([np.ones(N), -np.ones(N)])
    A_upp = sprand(N, n, density, rng)
    A_low = sprand(N, n, density, rng)
    Ad = sp.vstack([A_upp / np.sqrt(n) + (A_upp != 0.0).astype(float) / n,
                    A_low / np.sqrt(n) - (A_low != 0.0).astype(float) / n], format="csc")
    Im = sp.eye(m, format="csc")
    P = sp.block_diag([sp.eye(n), sp.csc_matrix((m, m))], format="csc")
    q = np.hstack([np.zeros(n), gamma * np.ones(m)])
    A = sp.bmat([[sp.diags(b) @ Ad, -Im], [None, Im]], format="csc")
    l = np.hstack([-np.inf * np.ones(m), np.zeros(m)])
    u = np.hstack([-np.ones(m), np.inf * np.ones(m)])
    return dict(P=P, q=q, A=A, l=l, u=u, name=f"svm_n{n}_m{m}")


_MPC_AD = np.array([
    [1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0., 0.],
    [0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0., 0.],
    [0., 0., 1., 0., 0., 0., 0., 0., 0.1, 0., 0., 0.],
    [0.0488, 0., 0., 1., 0., 0., 0.0016, 0., 0., 0.0992, 0., 0.],
    [0., -0.0488, 0., 0., 1., 0., 0., -0.0016, 0., 0., 0.0992, 0.],
    [0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0., 0.0992],
    [0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0., 0.],
    [0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0., 0.],
    [0., 0., 0., 0., 0., 0., 0., 0., 1., 0., 0., 0.],
    [0.9734, 0., 0., 0., 0., 0., 0.0488, 0., 0., 0.9846, 0., 0.],
    [0., -0.9734, 0., 0., 0., 0., 0., -0.0488, 0., 0., 0.9846, 0.],
    [0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0., 0.9846]])
_MPC_BD = np.array([
    [0., -0.0726, 0., 0.0726],
    [-0.0726, 0., 0.0726, 0.],
    [-0.0152, 0.0152, -0.0152, 0.0152],
    [-0., -0.0006, -0., 0.0006],
    [0.0006, 0., -0.0006, 0.0000],
    [0.0106, 0.0106, 0.0106, 0.0106],
    [0, -1.4512, 0., 1.4512],
    [-1.4512, 0., 1.4512, 0.],
    [-0.3049, 0.3049, -0.3049, 0.3049],
    [-0., -0.0236, 0., 0.0236],
    [0.0236, 0., -0.0236, 0.],
    [0.2107, 0.2107, 0.2107, 0.2107]])


def mpc(N=12, x0=None, seed=1):
    """docs/examples/mpc.rst:30-90 quadcopter; N=12 gives n=204, m=360 (BASELINE config 5 asks
    for n~200, m~400).  x0 defaults to a seeded U(-1,1)*0.1 perturbation."""
    rng = np.random.default_rng(seed)
    Ad, Bd = sp.csc_matrix(_MPC_AD), sp.csc_matrix(_MPC_BD)
    nx, nu = Bd.shape
    u0 = 10.5916
    umin = np.array([9.6, 9.6, 9.6, 9.6]) - u0
    umax = np.array([13., 13., 13., 13.]) - u0
    xmin = np.array([-np.pi / 6, -np.pi / 6, -np.inf, -np.inf, -np.inf, -1.] + [-np.inf] * 6)
    xmax = np.array([np.pi / 6, np.pi / 6] + [np.inf] * 10)
    Q = sp.diags([0., 0., 10., 10., 10., 10., 0., 0., 0., 5., 5., 5.])
    QN = Q
    R = 0.1 * sp.eye(4)
    if x0 is None:
        x0 = 0.1 * (2 * rng.random(nx) - 1)
    xr = np.array([0., 0., 1.] + [0.] * 9)
    P = sp.block_diag([sp.kron(sp.eye(N), Q), QN, sp.kron(sp.eye(N), R)], format="csc")
    q = np.hstack([np.kron(np.ones(N), -Q @ xr), -QN @ xr, np.zeros(N * nu)])
    Ax = sp.kron(sp.eye(N + 1), -sp.eye(nx)) + sp.kron(sp.eye(N + 1, k=-1), Ad)
    Bu = sp.kron(sp.vstack([sp.csc_matrix((1, N)), sp.eye(N)]), Bd)
    Aeq = sp.hstack([Ax, Bu])
    leq = np.hstack([-x0, np.zeros(N * nx)])
    Aineq = sp.eye((N + 1) * nx + N * nu)
    lineq = np.hstack([np.kron(np.ones(N + 1), xmin), np.kron(np.ones(N), umin)])
    uineq = np.hstack([np.kron(np.ones(N + 1), xmax), np.kron(np.ones(N), umax)])
    A = sp.vstack([Aeq, Aineq], format="csc")
    l = np.hstack([leq, lineq])
    u = np.hstack([leq, uineq])
    return dict(P=P, q=q, A=A, l=l, u=u, name=f"mpc_N{N}", nx=nx)


def kkt_bytes_per_cg_iter(n, m, nnzA, nnzP_full, F=8, I=4, rho_is_vec=False):
    """Algorithmic HBM bytes of one PCG iteration:
    K.p = SpMV(P) + SpMV(A) + SpMV(A') and the 8 n F of the vector phase."""
    kp = (nnzP_full + 2 * nnzA) * (F + I) + (2 * n + m + 3) * I + (6 * n + 2 * m) * F
    if rho_is_vec:
        kp += m * F
    return kp + 8 * n * F


def nnz_P_full(P):
    """Entries of the full symmetric CSR with structurally full diagonal (cuda_csr.cu:525)."""
    Pu = sp.triu(sp.csc_matrix(P), format="csc")
    ndiag = int((Pu.diagonal() != 0).sum())
    # structural diagonal count
    coo = Pu.tocoo()
    ndiag = int((coo.row == coo.col).sum())
    return 2 * (Pu.nnz - ndiag) + P.shape[0]


# ------------------------------------------------------------------------------------------------
# Block-seeded generators for the row-sharded mode (BASELINE configs[3]): the data matrix is defined
# as a stack of sample blocks, each drawn from its own seeded stream, so that the GLOBAL problem does
# not depend on the number of ranks and a rank can generate just the samples it owns -- no rank ever
# builds (or slices) the whole 1e8-nnz problem.  `*_shard(rank, world, ...)` returns the rank's QP in the
# column-split layout of osqp_b200.dist (columns = [features, shared by all ranks ; the slack columns
# of the rank's samples]), plus the global ids of its rows and columns; world = 1 gives the whole QP in
# the variable order of the docs example.

SAMPLE_BLOCK = 1 << 15


def _owned_blocks(n_samples, rank, world, block=SAMPLE_BLOCK):
    nb = -(-n_samples // block)
    j0, j1 = (rank * nb) // world, ((rank + 1) * nb) // world
    return [(j, j * block, min((j + 1) * block, n_samples)) for j in range(j0, j1)]


def _data_rows(n_features, blocks, density, seed, s0):
    """Entries of the rows [s0, s1) of the data matrix as (key-sorted) CSC pieces: returns (col, row -
    s0, value) sorted by column then row, duplicates removed, one independent stream per sample block."""
    keys, vals = [], []
    m_loc = blocks[-1][2] - s0 if blocks else 0
    for j, b0, b1 in blocks:
        rng = np.random.default_rng([seed, 7919, j])
        cnt = int(round(density * (b1 - b0) * n_features))
        rows = rng.integers(0, b1 - b0, size=cnt, dtype=np.int64) + (b0 - s0)
        cols = rng.integers(0, n_features, size=cnt, dtype=np.int64)
        key = np.unique(cols * m_loc + rows)
        keys.append(key)
        vals.append(rng.random(key.size))
    if not keys:
        return np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0)
    key, val = np.concatenate(keys), np.concatenate(vals)
    o = np.argsort(key, kind="stable")          # blocks cover disjoint rows: keys are unique
    key, val = key[o], val[o]
    return key // m_loc, key % m_loc, val


def _csc_with_slacks(n_features, m_loc, col, row, val, slack_blocks):
    """CSC of [B | S_1 | S_2 ...] stacked over row blocks: B = (col, row, val) sorted by column, and each
    slack block k is a list of (row_offset, value) pairs giving the entries of slack column i at rows
    row_offset + i.  Built directly from index arithmetic (no bmat / COO round trip)."""
    nnzB = col.size
    indptr = [np.concatenate([[0], np.cumsum(np.bincount(col, minlength=n_features))])]
    indices, data = [row], [val]
    base = nnzB
    for entries in slack_blocks:
        k = len(entries)
        ii = np.arange(m_loc, dtype=np.int64)
        idx = np.empty((m_loc, k), dtype=np.int64)
        dat = np.empty((m_loc, k))
        for t, (off, v) in enumerate(sorted(entries)):
            idx[:, t] = off + ii
            dat[:, t] = v
        indices.append(idx.ravel())
        data.append(dat.ravel())
        indptr.append(base + k * (ii + 1))
        base += k * m_loc
    nrows = max(off for entries in slack_blocks for off, _ in entries) + m_loc if slack_blocks else m_loc
    ncols = n_features + m_loc * len(slack_blocks)
    A = sp.csc_matrix((np.concatenate(data), np.concatenate(indices).astype(np.int32),
                       np.concatenate(indptr).astype(np.int32)), shape=(nrows, ncols))
    A.has_sorted_indices = True
    return A


def svm_shard(rank, world, n_features=10_000, n_samples=10_000_000, density=1e-3, gamma=1.0, seed=1,
              block=SAMPLE_BLOCK):
    """docs/examples/svm.rst:37-57, variables (x, t); samples [s0, s1) of rank `rank`."""
    n, m = n_features, 2 * (n_samples // 2)
    N = m // 2
    blocks = _owned_blocks(m, rank, world, block)
    s0, s1 = blocks[0][1], blocks[-1][2]
    mS = s1 - s0
    col, row, val = _data_rows(n, blocks, density, seed, s0)
    b = np.where(np.arange(s0, s1) < N, 1.0, -1.0)
    # Ad = A / sqrt(n) +- (A != 0) / n ; constraint rows diag(b) Ad
    val = b[row] * (val / np.sqrt(n) + b[row] / n)
    A = _csc_with_slacks(n, mS, col, row, val, [[(0, -1.0), (mS, 1.0)]])
    P = sp.block_diag([sp.eye(n), sp.csc_matrix((mS, mS))], format="csc")
    q = np.hstack([np.zeros(n), gamma * np.ones(mS)])
    l = np.hstack([-np.inf * np.ones(mS), np.zeros(mS)])
    u = np.hstack([-np.ones(mS), np.inf * np.ones(mS)])
    cols = np.concatenate([np.arange(n), n + np.arange(s0, s1)])
    rows = np.concatenate([np.arange(s0, s1), m + np.arange(s0, s1)])
    return dict(P=P, q=q, A=A, l=l, u=u, n_shared=n, cols=cols, rows=rows, n_global=n + m, m_global=2 * m,
                name=f"svm_n{n}_m{m}_shard{rank}of{world}")


def huber_shard(rank, world, n_features=10_000, n_samples=10_000_000, density=1e-3, seed=1, block=SAMPLE_BLOCK):
    """docs/examples/huber.rst:44-63, variables (x, u, r, s); samples [s0, s1) of rank `rank`."""
    n, m = n_features, n_samples
    blocks = _owned_blocks(m, rank, world, block)
    s0, s1 = blocks[0][1], blocks[-1][2]
    mS = s1 - s0
    col, row, val = _data_rows(n, blocks, density, seed, s0)
    x_true = np.random.default_rng([seed, 104729]).standard_normal(n) / np.sqrt(n)
    b = np.bincount(row, weights=val * x_true[col], minlength=mS)
    for j, b0, b1 in blocks:
        rng = np.random.default_rng([seed, 15485863, j])
        ind95 = (rng.random(b1 - b0) < 0.95).astype(float)
        b[b0 - s0:b1 - s0] += 0.5 * rng.standard_normal(b1 - b0) * ind95 + 10.0 * rng.random(b1 - b0) * (1.0 - ind95)
    # A = [Ad -I -I I; 0 0 I 0; 0 0 0 I] restricted to the rank's samples
    A = _csc_with_slacks(n, mS, col, row, val, [[(0, -1.0)], [(0, -1.0), (mS, 1.0)], [(0, 1.0), (2 * mS, 1.0)]])
    P = sp.block_diag([sp.csc_matrix((n, n)), 2 * sp.eye(mS), sp.csc_matrix((2 * mS, 2 * mS))], format="csc")
    q = np.append(np.zeros(mS + n), 2 * np.ones(2 * mS))
    l = np.hstack([b, np.zeros(2 * mS)])
    u = np.hstack([b, np.inf * np.ones(2 * mS)])
    S = np.arange(s0, s1)
    cols = np.concatenate([np.arange(n), n + S, n + m + S, n + 2 * m + S])
    rows = np.concatenate([S, m + S, 2 * m + S])
    return dict(P=P, q=q, A=A, l=l, u=u, n_shared=n, cols=cols, rows=rows, n_global=n + 3 * m, m_global=3 * m,
                name=f"huber_n{n}_m{m}_shard{rank}of{world}")


def assemble_shards(shards):
    """The global QP from the shards of all ranks (tests / small sizes only)."""
    n, m = shards[0]["n_global"], shards[0]["m_global"]
    ns = shards[0]["n_shared"]
    A = sp.lil_matrix((m, n))
    P = sp.lil_matrix((n, n))
    q, l, u = np.zeros(n), np.zeros(m), np.zeros(m)
    for sh in shards:
        Ac = sp.coo_matrix(sh["A"])
        A = A + sp.coo_matrix((Ac.data, (sh["rows"][Ac.row], sh["cols"][Ac.col])), shape=(m, n))
        Pc = sp.coo_matrix(sh["P"])
        own = (Pc.row >= ns) | (Pc.col >= ns)
        first = sh is shards[0]
        keep = own | first                          # the shared block of P is the same on every rank
        P = P + sp.coo_matrix((Pc.data[keep], (sh["cols"][Pc.row[keep]], sh["cols"][Pc.col[keep]])), shape=(n, n))
        q[sh["cols"]] = sh["q"]
        l[sh["rows"]], u[sh["rows"]] = sh["l"], sh["u"]
    return dict(P=sp.csc_matrix(P), q=q, A=sp.csc_matrix(A), l=l, u=u)
