"""CPU restatements (numpy) of the three algorithms introduced behind the C-ABI this round, checked
against their textbook forms.  They pin the MATH the kernels implement; the kernels themselves are
compared with these properties on the GPU (tests/test_gpu_kernels.py).

  * the CG recurrence of the graph PCG driver: alpha AND beta from three dots of the operator pass
    (osqp_b200/csrc/pcg_graph.cu cg_step_scalars, g_update_fused) vs textbook Jacobi-PCG
    (reference: algebra/cuda/lin_sys/indirect/cuda_pcg.cu:113-208);
  * the device transpose (osqp_b200/csrc/transpose.cu): scatter through an atomic cursor in ANY
    order + per-row rank sort == stable counting sort;
  * the carried product A x of the fused x/z/y update (src/auxil.c:172-184 relaxation step).
"""
import numpy as np
import pytest
import scipy.sparse as sp


def _kkt(n, m, seed):
    rng = np.random.default_rng(seed)
    A = sp.random(m, n, density=0.15, format="csr", random_state=seed, data_rvs=lambda s: rng.standard_normal(s))
    M = sp.random(n, n, density=0.1, format="csr", random_state=seed + 1, data_rvs=lambda s: rng.standard_normal(s))
    P = (M @ M.T).tocsr()
    rho = rng.uniform(0.05, 2.0, m)
    K = (P + 1e-6 * sp.eye(n) + A.T @ sp.diags(rho) @ A).tocsr()
    return K, rng.standard_normal(n)


def _pcg_textbook(K, b, minv, x0, eps, max_iter):
    x = x0.copy()
    r = K @ x - b
    y = minv * r
    p = -y
    rTy = r @ y
    it = 0
    hist = [x.copy()]
    while np.abs(r).max() > eps and it < max_iter:
        Kp = K @ p
        alpha = rTy / (p @ Kp)
        x += alpha * p
        r += alpha * Kp
        y = minv * r
        rTy_new = r @ y
        beta = rTy_new / rTy
        rTy = rTy_new
        p = beta * p - y
        it += 1
        hist.append(x.copy())
    return x, it, hist


def _pcg_three_dots(K, b, minv, x0, eps, max_iter):
    """loop body = operator pass (Kp + 3 dots) then ONE vector update that already knows beta"""
    x = x0.copy()
    r = K @ x - b
    p = -(minv * r)
    rTy = r @ (minv * r)
    it = 0
    hist = [x.copy()]
    while np.abs(r).max() > eps and it < max_iter:
        Kp = K @ p
        pKp, rkp, kpkp = p @ Kp, r @ (minv * Kp), Kp @ (minv * Kp)
        alpha = rTy / pKp
        pred = rTy + alpha * (2.0 * rkp + alpha * kpkp)       # r+' M^-1 r+ without forming r+
        beta = max(pred / rTy, 0.0)
        x += alpha * p
        r += alpha * Kp
        y = minv * r
        p = beta * p - y
        rTy = r @ y                                            # the exact value replaces the prediction
        it += 1
        hist.append(x.copy())
    return x, it, hist


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_three_dot_cg_recurrence_matches_textbook_pcg(seed):
    K, b = _kkt(120, 200, seed)
    minv = 1.0 / K.diagonal()
    x0 = np.zeros(120)
    eps = 1e-9 * np.abs(b).max()
    xa, ita, ha = _pcg_textbook(K, b, minv, x0, eps, 500)
    xb, itb, hb = _pcg_three_dots(K, b, minv, x0, eps, 500)
    assert abs(ita - itb) <= 1 and ita > 5      # the stopping test sits at rounding level for one iterate
    scale = np.abs(xa).max()
    for u, v in zip(ha, hb):                                   # same iterates, not just the same limit
        assert np.abs(u - v).max() <= 1e-7 * scale
    assert np.abs(K @ xb - b).max() <= eps and np.abs(K @ xa - b).max() <= eps


def _transpose_device_algorithm(rp, ci, vx, nrows, ncols, rng):
    """count -> scan -> scatter in a RANDOM order (the order atomics happen to be served in) ->
    per-row rank sort by (column, source position)."""
    nnz = rp[-1]
    cnt = np.bincount(ci, minlength=ncols)
    rpt = np.concatenate([[0], np.cumsum(cnt)])
    src_row = np.repeat(np.arange(nrows), np.diff(rp))
    cursor = rpt[:-1].copy()
    tcol = np.empty(nnz, dtype=np.int64)
    tsrc = np.empty(nnz, dtype=np.int64)
    for k in rng.permutation(nnz):
        pos = cursor[ci[k]]
        cursor[ci[k]] += 1
        tcol[pos], tsrc[pos] = src_row[k], k
    out_c = np.empty(nnz, dtype=np.int64)
    out_v = np.empty(nnz)
    mp = np.empty(nnz, dtype=np.int64)
    for i in range(ncols):
        s, e = rpt[i], rpt[i + 1]
        c, sidx = tcol[s:e], tsrc[s:e]
        rank = np.array([np.sum((c < c[t]) | ((c == c[t]) & (sidx < sidx[t]))) for t in range(e - s)], dtype=np.int64)
        out_c[s + rank] = c
        out_v[s + rank] = vx[sidx]
        mp[sidx] = s + rank
    return rpt, out_c, out_v, mp


def test_rank_sorted_scatter_is_a_stable_counting_sort_whatever_the_atomic_order():
    rng = np.random.default_rng(4)
    M = sp.random(40, 25, density=0.2, format="csr", random_state=4, data_rvs=lambda s: rng.standard_normal(s))
    M.sort_indices()
    ref = sp.csr_matrix(M.T)
    ref.sort_indices()
    outs = [_transpose_device_algorithm(M.indptr, M.indices, M.data, 40, 25, np.random.default_rng(sd)) for sd in (0, 1)]
    for rpt, c, v, mp in outs:
        assert np.array_equal(rpt, ref.indptr) and np.array_equal(c, ref.indices) and np.array_equal(v, ref.data)
        assert np.array_equal(v[mp], M.data)
    assert all(np.array_equal(a, b) for a, b in zip(outs[0], outs[1]))


def test_carried_Ax_follows_the_relaxation_step():
    """x+ = alpha x~ + (1 - alpha) x and z~ = A x~  =>  A x+ = alpha z~ + (1 - alpha) A x."""
    rng = np.random.default_rng(6)
    A = sp.random(80, 30, density=0.2, format="csr", random_state=6)
    x = rng.standard_normal(30)
    Ax = A @ x
    alpha = 1.6
    for _ in range(200):
        xt = rng.standard_normal(30)
        zt = A @ xt
        x = alpha * xt + (1 - alpha) * x
        Ax = alpha * zt + (1 - alpha) * Ax
    assert np.abs(Ax - A @ x).max() <= 1e-12 * max(1.0, np.abs(Ax).max())


def test_column_split_cg_equals_the_global_cg():
    """The row-sharded CG of pcg_graph.cu (b200_pcg_sharded_solve) emulated with numpy for 3 virtual
    ranks: every rank applies (P + sigma I) only to the columns it counts (rank 0: shared + own, the
    others: own) plus A_r' rho A_r, the SHARED head of the partial is summed over the ranks, and the
    dots are summed over [shared once ; every rank's own columns].  Must reproduce the CG on the whole
    KKT operator iterate for iterate."""
    from osqp_b200 import problems
    from osqp_b200.dist import plan_column_split, shard_problem_split, assemble_solution
    world, sigma = 3, 1e-6
    pb = problems.huber(12, 240, density=0.2, seed=8)
    A = sp.csr_matrix(pb["A"])
    P = sp.csc_matrix(pb["P"])
    Pfull = (sp.triu(P) + sp.triu(P, 1).T).tocsr()
    m, n = A.shape
    rng = np.random.default_rng(8)
    rho_g = rng.uniform(0.1, 1.0, m)
    b_g = rng.standard_normal(n)
    K = (Pfull + sigma * sp.eye(n) + A.T @ sp.diags(rho_g) @ A).tocsr()
    minv_g = 1.0 / K.diagonal()
    x_ref, it_ref, hist_ref = _pcg_three_dots(K, b_g, minv_g, np.zeros(n), 1e-10, 60)

    plan = plan_column_split(P, A, world)
    ns = plan["shared"].size
    assert 0 < ns < n
    ranks = []
    for r in range(world):
        sh = shard_problem_split(pb, r, plan)
        cols, rows = plan["cols"][r], plan["rows"][r]
        Ar = sp.csr_matrix(sh["A"])[:len(rows)]
        Pr = sp.csc_matrix(sh["P"])
        Pr = (sp.triu(Pr) + sp.triu(Pr, 1).T).tocsr()
        counted = np.ones(cols.size, dtype=bool)
        if r > 0:
            counted[:ns] = False                    # shared columns: P + sigma I carried by rank 0 only
        ranks.append(dict(A=Ar, P=Pr, rho=rho_g[rows], cols=cols, counted=counted, b=b_g[cols],
                          minv=minv_g[cols], x=np.zeros(cols.size)))

    def apply_K(vecs):
        parts = []
        for rk, v in zip(ranks, vecs):
            kp = rk["A"].T @ (rk["rho"] * (rk["A"] @ v))
            kp += np.where(rk["counted"], rk["P"] @ v + sigma * v, 0.0)
            parts.append(kp)
        head = sum(p[:ns] for p in parts)           # the ONE vector exchange: all-reduce of the shared head
        for p in parts:
            p[:ns] = head
        return parts

    def dot(us, vs):                                # shared slice once (rank 0) + every rank's own columns
        return sum(float(u[(0 if i == 0 else ns):] @ v[(0 if i == 0 else ns):]) for i, (u, v) in enumerate(zip(us, vs)))

    xs = [rk["x"] for rk in ranks]
    rs = [kp - rk["b"] for kp, rk in zip(apply_K(xs), ranks)]
    ps = [-(rk["minv"] * r) for rk, r in zip(ranks, rs)]
    rTy = dot(rs, [rk["minv"] * r for rk, r in zip(ranks, rs)])
    for it in range(it_ref):
        Kps = apply_K(ps)
        yk = [rk["minv"] * kp for rk, kp in zip(ranks, Kps)]
        pKp, rkp, kpkp = dot(ps, Kps), dot(rs, yk), dot(Kps, yk)
        alpha = rTy / pKp
        beta = max((rTy + alpha * (2 * rkp + alpha * kpkp)) / rTy, 0.0)
        xs = [x + alpha * p for x, p in zip(xs, ps)]
        rs = [r + alpha * kp for r, kp in zip(rs, Kps)]
        ys = [rk["minv"] * r for rk, r in zip(ranks, rs)]
        ps = [beta * p - y for p, y in zip(ps, ys)]
        rTy = dot(rs, ys)
        xg, _ = assemble_solution([(x, np.zeros(len(plan["rows"][i]))) for i, x in enumerate(xs)], n, m, plan=plan)
        # same iterates; summation orders differ, and CG amplifies rounding slowly with the iteration count
        tol = (1e-11 if it < 10 else 1e-6) * max(1.0, np.abs(hist_ref[it + 1]).max())
        assert np.abs(xg - hist_ref[it + 1]).max() <= tol
        for x in xs[1:]:                            # the replicated slice stays identical on every rank
            assert np.array_equal(x[:ns], xs[0][:ns])
