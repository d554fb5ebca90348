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
