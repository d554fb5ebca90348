"""CPU tests of the N > 1 paths (gloo, world_size 2): row partitioning of the sharded mode, the
unique-id hand-off used to build the NCCL communicator, and the rank bookkeeping of bench.py."""
import os
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import ROOT
from osqp_b200 import problems
from osqp_b200.dist import partition_rows, shard_problem


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_partition_covers_rows_and_balances_nonzeros(world):
    pb = problems.lasso(40, 400, density=0.1, seed=2)
    A = sp.csr_matrix(pb["A"])
    n = A.shape[1]
    b = partition_rows(A, world)
    assert b[0] == 0 and b[-1] == A.shape[0] and (np.diff(b) > 0).all()
    if world > 1:      # only the sharded mode needs block lengths to differ from n
        assert not (np.diff(b) == n).any()
    nnz = np.array([A[b[r]:b[r + 1]].nnz for r in range(world)])
    assert nnz.max() <= 1.6 * nnz.mean() + n
    # the shards reassemble to the original problem
    parts = [shard_problem(pb, r, world) for r in range(world)]
    assert (sp.vstack([p["A"] for p in parts]) != sp.csc_matrix(pb["A"])).nnz == 0
    assert np.array_equal(np.concatenate([p["l"] for p in parts]), pb["l"])
    assert np.array_equal(np.concatenate([p["u"] for p in parts]), pb["u"])


def test_partition_avoids_blocks_of_exactly_n_rows():
    A = sp.random(20, 10, density=0.3, format="csr", random_state=0)   # m = 2n: the even split hits n
    b = partition_rows(A, 2)
    assert not (np.diff(b) == 10).any() and b[-1] == 20


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, torch, torch.distributed as dist
    from osqp_b200 import problems
    from osqp_b200.dist import exchange_unique_id, shard_problem
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = exchange_unique_id(lambda: bytes(range(128)), dist)
    assert uid == bytes(range(128)), "unique id not identical on every rank"
    pb = problems.portfolio(300, 20, density=0.2, seed=1)
    sh = shard_problem(pb, rank, world)
    rows = torch.tensor([sh["A"].shape[0], sh["A"].nnz], dtype=torch.int64)
    dist.all_reduce(rows)
    assert rows[0].item() == pb["A"].shape[0] and rows[1].item() == pb["A"].nnz
    # max-over-ranks timing reduction used by bench.py
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t.item() == world
    dist.barrier()
    if rank == 0:
        print("WORKER_OK")
    dist.destroy_process_group()
""")


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert "WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_bench_reference_arm_other_ranks_exit_quietly():
    """--impl reference under torchrun: only rank 0 works and prints"""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("family", ["lasso", "huber", "svm", "portfolio"])
def test_column_split_plan_partitions_the_problem(family, world):
    """Column-split layout: every row and every column has exactly one home, a rank's
    rows touch only [shared ; its own] columns, P never couples columns of different ranks, and for
    the epigraph families only the feature columns are shared."""
    from osqp_b200.dist import plan_column_split, shard_problem_split
    pb = {"lasso": lambda: problems.lasso(50, 500, density=0.05, seed=2),
          "huber": lambda: problems.huber(20, 400, density=0.1),
          "svm": lambda: problems.svm(20, 400, density=0.1),
          "portfolio": lambda: problems.portfolio(300, 20, density=0.2)}[family]()
    A = sp.csr_matrix(pb["A"])
    P = sp.csc_matrix(pb["P"])
    Pfull = (sp.triu(P) + sp.triu(P, 1).T).tocsr()
    m, n = A.shape
    plan = plan_column_split(P, A, world)
    ns = plan["shared"].size
    owned_cols, owned_rows = np.zeros(n, dtype=int), np.zeros(m, dtype=int)
    owned_cols[plan["shared"]] += 1
    nnzA = nnzP_local = 0
    for r in range(world):
        R, Cc = plan["rows"][r], plan["cols"][r]
        assert np.array_equal(Cc[:ns], plan["shared"])
        owned_rows[R] += 1
        owned_cols[Cc[ns:]] += 1
        sh = shard_problem_split(pb, r, plan)
        assert sh["A"].shape[0] != sh["A"].shape[1]            # lengths tell row from column vectors
        assert A[R].nnz == sh["A"].nnz                          # no entry of these rows lies outside Cc
        nnzA += sh["A"].nnz
        assert np.array_equal(sh["q"], np.asarray(pb["q"])[Cc])
        # P restricted to this rank's columns keeps every coupling of its owned columns
        own = Cc[ns:]
        if own.size:
            assert Pfull[own].nnz == Pfull[own][:, Cc].nnz
            nnzP_local += Pfull[own][:, own].nnz
    assert (owned_rows == 1).all() and (owned_cols == 1).all() and nnzA == A.nnz
    if family != "portfolio":
        assert ns <= 50 and ns < 0.2 * n
    else:
        assert ns == n                                          # no low-degree cut: falls back to row blocks


SPLIT_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import numpy as np, scipy.sparse as sp, torch, torch.distributed as dist
    from osqp_b200 import problems
    from osqp_b200.dist import plan_column_split, shard_problem_split, assemble_solution
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    pb = problems.huber(15, 300, density=0.15, seed=4)
    A = sp.csr_matrix(pb["A"]); n, m = A.shape[1], A.shape[0]
    plan = plan_column_split(pb["P"], pb["A"], world)
    sh = shard_problem_split(pb, rank, plan)
    ns = plan["shared"].size
    rng = np.random.default_rng(11)                       # same "solution" on every rank
    x_true, y_true = rng.standard_normal(n), rng.standard_normal(m)
    rows, cols = plan["rows"][rank], plan["cols"][rank]
    x_loc = x_true[cols]
    y_loc = np.concatenate([y_true[rows], np.zeros(sh["padded"])])
    # rule 1 (algebra/b200/vector.c DIST_REDUCE): a reduction over a column vector counts the shared
    # slice on rank 0 only, then the ranks are summed / maxed
    off = 0 if rank == 0 else ns
    t = torch.tensor([float(sh["q"][off:] @ x_loc[off:])], dtype=torch.float64)
    dist.all_reduce(t)
    assert abs(t.item() - float(pb["q"] @ x_true)) < 1e-9 * max(1.0, abs(float(pb["q"] @ x_true)))
    mx = torch.tensor([float(np.abs(x_loc[off:]).max()) if x_loc[off:].size else 0.0], dtype=torch.float64)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    assert mx.item() == float(np.abs(x_true).max())
    # rule 2 (matrix.c Atxpy / pcg_graph.cu exchange_vector): A_r' y_r is complete on the owned columns
    # and needs a sum over the ranks on the shared head only
    part = sh["A"].T @ y_loc
    head = torch.from_numpy(part[:ns].copy())
    dist.all_reduce(head)
    full = A.T @ y_true
    assert np.allclose(head.numpy(), full[plan["shared"]], rtol=1e-12, atol=1e-12)
    assert np.allclose(part[ns:], full[cols[ns:]], rtol=1e-12, atol=1e-12)
    # rule 3: A_r x_r needs no exchange at all
    assert np.allclose((sh["A"] @ x_loc)[:len(rows)], (A @ x_true)[rows], rtol=1e-12, atol=1e-12)
    # assembly of the global solution from the per-rank slices
    parts = [None] * world
    dist.all_gather_object(parts, (x_loc, y_loc))
    x, y = assemble_solution(parts, n, m, plan=plan)
    assert np.array_equal(x, x_true) and np.array_equal(y, y_true)
    dist.barrier()
    if rank == 0:
        print("SPLIT_WORKER_OK", ns, n)
    dist.destroy_process_group()
""")


def test_gloo_column_split_layout_rules(tmp_path):
    """world_size 2 on CPU: the three exchange rules of the column-split layout (reductions count the
    shared slice once; A'y is exchanged on the shared head only; A x needs nothing) and the assembly
    of the global solution."""
    script = tmp_path / "split_worker.py"
    script.write_text(SPLIT_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert "SPLIT_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("family", ["lasso", "huber", "svm", "portfolio"])
def test_csc_native_slicer_equals_scipy_slicing(family, world):
    """The shard builder works on the CSC arrays directly (no CSC -> CSR conversion of the whole
    matrix on every rank); it must produce exactly what scipy's fancy indexing produces."""
    from osqp_b200.dist import plan_column_split, shard_problem_split
    pb = {"lasso": lambda: problems.lasso(60, 700, density=0.04, seed=5),
          "huber": lambda: problems.huber(25, 500, density=0.1, seed=5),
          "svm": lambda: problems.svm(25, 500, density=0.1, seed=5),
          "portfolio": lambda: problems.portfolio(400, 25, density=0.2, seed=5)}[family]()
    plan = plan_column_split(pb["P"], pb["A"], world)
    Acsr = sp.csr_matrix(pb["A"])
    for r in range(world):
        fast = shard_problem_split(pb, r, plan)
        ref = shard_problem_split(pb, r, plan, Acsr=Acsr)
        for key in ("A", "P"):
            a, b = sp.csc_matrix(fast[key]), sp.csc_matrix(ref[key])
            a.sort_indices(); b.sort_indices()
            assert a.shape == b.shape and np.array_equal(a.indptr, b.indptr)
            assert np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data)
        for key in ("q", "l", "u"):
            assert np.array_equal(fast[key], ref[key])
        assert fast["padded"] == ref["padded"] and fast["n_shared"] == ref["n_shared"]


def test_bench_reference_arm_line_contract():
    """`bench.py --impl reference` (the reference's CPU path timed on the host cores): ONE JSON line on
    stdout with the keys the driver reads, same metric / unit as the B200 arm."""
    import json
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "admm_iters_per_sec" and d["unit"] == "iter/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["unit"] == d["unit"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("family", ["svm", "huber"])
def test_block_seeded_shards_reassemble_to_the_world_size_1_problem(family, world):
    """bench.py's sharded workload: the global QP is defined by seeded sample blocks, so the shards of
    any world size put together are exactly the problem a single rank generates, every shard is the
    [rows of the rank] x [shared ; owned] slice of it, and only the features are shared."""
    gen = {"svm": problems.svm_shard, "huber": problems.huber_shard}[family]
    kw = dict(n_features=40, n_samples=3000, density=0.05, seed=1, block=256)
    whole = gen(0, 1, **kw)
    shards = [gen(r, world, **kw) for r in range(world)]
    g = problems.assemble_shards(shards)
    for key in ("P", "A"):
        d = sp.csc_matrix(whole[key]) - g[key]
        assert d.nnz == 0 or abs(d).max() == 0
    for key in ("q", "l", "u"):
        assert np.array_equal(whole[key], g[key])
    Aw = sp.csc_matrix(whole["A"]).tocsr()
    seen_rows = np.concatenate([s["rows"] for s in shards])
    assert np.array_equal(np.sort(seen_rows), np.arange(whole["A"].shape[0]))          # every row has one home
    owned = np.concatenate([s["cols"][s["n_shared"]:] for s in shards])
    assert np.array_equal(np.sort(np.concatenate([shards[0]["cols"][:shards[0]["n_shared"]], owned])),
                          np.arange(whole["A"].shape[1]))                               # so has every column
    for s in shards:
        assert s["n_shared"] == 40 and s["A"].shape[0] != s["A"].shape[1]
        sub = Aw[s["rows"]][:, s["cols"]]
        assert abs(sub - s["A"]).max() == 0
        # a rank's rows touch no column owned by another rank
        mask = np.ones(whole["A"].shape[1], dtype=bool)
        mask[s["cols"]] = False
        assert Aw[s["rows"]][:, np.nonzero(mask)[0]].nnz == 0


PEER_WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r})
    import torch.distributed as dist
    from osqp_b200.devmem import kernels
    from osqp_b200.dist import ShardedOSQP, enable_peer_exchange
    dist.init_process_group("gloo")
    k = kernels("f64")
    # no CUDA device here: the export fails on every rank, the handshake must end in a consistent "no"
    assert enable_peer_exchange(k, dist) is False
    os.environ["B200_DIST_NO_P2P"] = "1"
    assert enable_peer_exchange(k, dist) is False
    for bad in (dict(polishing=1), dict(time_limit=1.0), dict(adaptive_rho=2)):
        try:
            ShardedOSQP._check_settings(bad)
        except ValueError:
            continue
        raise AssertionError(f"{{bad}} accepted in the row-sharded mode")
    ShardedOSQP._check_settings(dict(adaptive_rho=1, eps_abs=1e-3))
    dist.barrier()
    if dist.get_rank() == 0:
        print("PEER_WORKER_OK")
    dist.destroy_process_group()
""")


def test_gloo_peer_exchange_handshake_and_settings_guard(tmp_path):
    script = tmp_path / "peer_worker.py"
    script.write_text(PEER_WORKER.format(root=str(ROOT)))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    env.pop("B200_DIST_NO_P2P", None)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29537", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert "PEER_WORKER_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_bench_reference_arm_follows_the_arm_workload():
    """under torchrun the B200 arm solves the sharded SVM: the reference arm must name the same workload"""
    import json
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert d["impl"] == "reference" and d["scaling"] == "strong" and "svm" in d["config"]["workload"]
    assert d["cpu_baseline"]["cores"] == 1 and d["e2e"]["value"] == d["value"]
