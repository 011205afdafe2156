"""Clause-range sharded mode: same trajectory and result as the unsharded solve."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "tests"))
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat
from alllsatisfiabilitysolver_b200.sharded import ShardedSolver, partition
from shard_oracle_backend import OracleShardBackend
from oracle.oracle import Oracle

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n, k, d, seed = 3000, 5, 4, 9
lits = bounded_degree_ksat(n, k, d, seed=77)
m = lits.shape[0]
ss = ShardedSolver(OracleShardBackend(m), rank, world)
ss.upload_full(n, lits)
ss.backend.randomize(seed)
st = ss.solve(seed)
mine = torch.from_numpy(ss.backend.get_assignment().copy())
gathered = [torch.empty_like(mine) for _ in range(world)]
dist.all_gather(gathered, mine)
same = all(bool((g == mine).all()) for g in gathered)
o = Oracle()
off = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
v = o.randomize(n, seed)
so, tu, ts = o.solve(n, off, lits.reshape(-1), v, seed, trace=True)
ok = (same and st.status == 0 and (st.n_iterations, st.n_resamples, st.sum_mis_size, st.avg_mis_size) ==
      (so.n_iterations, so.n_resamples, so.sum_mis_size, so.avg_mis_size) and st.trace_u == [int(x) for x in tu]
      and st.trace_s == [int(x) for x in ts] and bool((mine.numpy() == v).all()) and o.verify(off, lits.reshape(-1), v)
      and st.n_clause_evals == m * so.n_iterations)
lo, hi = partition(m, world)[rank]
print(json.dumps(dict(rank=rank, ok=bool(ok), iters=st.n_iterations, range=[lo, hi])), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 1)
'''


def test_partition_is_contiguous_and_balanced():
    from alllsatisfiabilitysolver_b200.sharded import partition

    for m in (0, 1, 7, 1000, 39_999_887):
        for w in (1, 2, 3, 8):
            p = partition(m, w)
            assert p[0][0] == 0 and p[-1][1] == m and all(p[i][1] == p[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in p]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_host_logic_under_gloo(world, tmp_path, oracle):
    """world_size>1 on CPU: the driver's partition / all-gather / termination logic with the oracle as the per-rank
    compute.  Every rank must end with the identical assignment and the unsharded oracle trajectory."""
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), str(script)],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count('"ok": true') == world


@pytest.mark.gpu
@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sharded_equals_single_gpu_bit_exact(n_shards, oracle):
    """SURVEY test plan item 6: R shards simulated as R handles on one GPU give the single-GPU assignment,
    violated/independent set sizes and statistics for the same seed (the exchange is a plain concatenation here)."""
    import torch

    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat
    from alllsatisfiabilitysolver_b200.sharded import CudaShardBackend, partition

    n, k, seed = 400_000, 8, 21          # 50 KB of assignment bits per handle; ~1.6 M clauses in total
    lits = bounded_degree_ksat(n, k, 32, seed=5)
    m = lits.shape[0]
    off = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
    v = oracle.randomize(n, seed)
    so, tu, ts = oracle.solve(n, off, lits.reshape(-1), v, seed, trace=True)

    backends = []
    for lo, hi in partition(m, n_shards):
        b = CudaShardBackend(0)
        b.upload(n, lits[lo:hi], lo)
        b.randomize(seed)
        backends.append(b)
    trace_u, trace_s = [], []
    for rnd in range(200):
        sends = [b.sweep_export() for b in backends]
        counts = [c for _, c in sends]
        cap = max(max(counts), 1)
        recs = torch.zeros((n_shards, cap, k + 1), dtype=torch.int32, device="cuda")
        for r, (t, c) in enumerate(sends):
            recs[r, :c] = t[:c]
        torch.cuda.synchronize()
        outs = [b.shard_round(recs, counts, seed, rnd) for b in backends]
        assert all(o == outs[0] for o in outs)
        trace_u.append(outs[0][0])
        trace_s.append(outs[0][1])
        if outs[0][0] == 0:
            break
    assert trace_u == [int(x) for x in tu] and trace_s == [int(x) for x in ts]
    for b in backends:
        assert np.array_equal(b.get_assignment(), v)
        st = b.stats()
        assert (st.n_iterations, st.n_resamples, st.sum_mis_size, st.avg_mis_size) == \
               (so.n_iterations, so.n_resamples, so.sum_mis_size, so.avg_mis_size)
        b.solver.close()


@pytest.mark.gpu
def test_sharded_nccl_two_gpus(tmp_path):
    """Real exchange over NCCL when the box has >= 2 GPUs (skipped on single-GPU boxes)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29655",
                        os.path.join(ROOT, "tools", "run_sharded.py"), "--scale", "0.02", "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.gpu
def test_sharded_p2p_fused_exchange_two_gpus():
    """Exchange fused into the kernels (NVLink P2P stores + arrival flags, CUDA IPC): needs >= 2 GPUs in one box."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29656",
                        os.path.join(ROOT, "tools", "run_sharded.py"), "--p2p", "--scale", "0.05", "--solves", "3", "--check"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.gpu
def test_sharded_p2p_persistent_kernels_two_gpus():
    """Same, with every rank's solve as one persistent kernel (exchange published after the grid barrier)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29657",
                        os.path.join(ROOT, "tools", "run_sharded.py"), "--p2p", "--persistent", "--scale", "0.05", "--solves", "3",
                        "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.gpu
def test_sharded_p2p_persistent_incremental_two_gpus():
    """Persistent kernels + incremental re-evaluation of each rank's own clause range (the violated records of an
    incremental round go through the same fused exchange)."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29658",
                        os.path.join(ROOT, "tools", "run_sharded.py"), "--p2p", "--persistent", "--incremental", "--scale", "0.05",
                        "--solves", "3", "--check"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.gpu
def test_p2p_persistent_incremental_single_rank_matches_oracle(oracle):
    """world = 1 runs the incremental branch of the sharded persistent kernel on one GPU: S comes from the previous
    round's records in the exchange region, the new violated clauses leave as records; trajectory == oracle."""
    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat
    from alllsatisfiabilitysolver_b200.sharded import P2PShardedSolver

    for n, k, d in ((300_000, 8, 32), (120_000, 7, 28), (2_000_000, 8, 32)):      # resident / resident / bucketed layout
        lits = bounded_degree_ksat(n, k, d, seed=16)
        m = lits.shape[0]
        off = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
        ss = P2PShardedSolver(0, 0, 1, persistent=True, flags=capi.FLAG_INCREMENTAL)
        ss.upload_range(n, lits, m, 0)
        used = 0
        for sd in (7, 8):
            ss.randomize(sd)
            st = ss.solve(sd)
            v = oracle.randomize(n, sd)
            so = oracle.solve(n, off, lits.reshape(-1), v, sd)
            assert (st.n_iterations, st.n_resamples, st.sum_mis_size, st.status) == (so.n_iterations, so.n_resamples, so.sum_mis_size, 0)
            assert np.array_equal(ss.get_assignment(), v) and ss.solver.verify()
            used += st.n_incremental_rounds
            if st.n_incremental_rounds:
                assert st.n_clause_evals < m * st.n_iterations
        assert used > 0
        ss.solver.close()


@pytest.mark.gpu
@pytest.mark.parametrize("persistent", [False, True], ids=["kernel_per_phase", "persistent_kernel"])
def test_p2p_mode_single_rank_matches_plain_solve(oracle, persistent):
    """world = 1 exercises the whole P2P code path (record export, flags, MIS over records) on one GPU."""
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat
    from alllsatisfiabilitysolver_b200.sharded import P2PShardedSolver

    n, k, seed = 300_000, 8, 4
    lits = bounded_degree_ksat(n, k, 32, seed=6)
    m = lits.shape[0]
    off = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
    ss = P2PShardedSolver(0, 0, 1, persistent=persistent)
    ss.upload_range(n, lits, m, 0)
    for sd in (seed, seed + 1):
        ss.randomize(sd)
        st = ss.solve(sd)
        v = oracle.randomize(n, sd)
        so = oracle.solve(n, off, lits.reshape(-1), v, sd)
        assert (st.n_iterations, st.n_resamples, st.sum_mis_size, st.status) == (so.n_iterations, so.n_resamples, so.sum_mis_size, 0)
        assert np.array_equal(ss.get_assignment(), v)
    # capacity overflow is reported, not hidden: 16 records cannot hold round 0
    tiny = P2PShardedSolver(0, 0, 1, persistent=persistent)
    tiny.upload_range(n, lits, m, 0, cap_records=16)
    tiny.randomize(1)
    with pytest.raises(tiny.capi.AlllError) as e:
        tiny.solve(1)
    assert e.value.status == tiny.capi.CAPACITY
    ss.solver.close()
    tiny.solver.close()
