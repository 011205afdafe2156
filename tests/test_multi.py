"""Several GPUs behind one call from one process (include/alll_b200.h: alll_multi_*).

The device list may name the same GPU several times (simulated ranks): on a single-GPU box this drives the whole fused
exchange of the clause-range sharded solve -- records stored into every rank's exchange region by the sweep, per-round
arrival flags, every rank deciding on the gathered violated set -- at world 2 and 3 against the oracle's trajectory.
With >= 2 GPUs the same tests also run on distinct devices (persistent kernels, peer access over NVLink)."""
import os
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from alllsatisfiabilitysolver_b200 import capi as m

    m.load()
    return m


def device_lists():
    import torch

    n = torch.cuda.device_count()
    out = [([0, 0], "two_ranks_one_gpu"), ([0, 0, 0], "three_ranks_one_gpu")]
    if n >= 2:
        out.append(([0, 1], "two_gpus"))
    if n >= 4:
        out.append(([0, 1, 2, 3], "four_gpus"))
    return out


def _instance(n=300_000, k=8, d=32, seed=6):
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

    lits = bounded_degree_ksat(n, k, d, seed=seed)
    m = lits.shape[0]
    return n, k, m, lits, np.arange(m + 1, dtype=np.uint64) * np.uint64(k)


@pytest.mark.parametrize("flags", [0, 4], ids=["full_sweeps", "incremental"])
def test_sharded_solve_in_one_process_matches_oracle(capi, oracle, flags):
    """alll_multi_solve == the oracle's solve for the same seed: Statistics, final assignment, validity -- for every
    device list this box offers, twice per handle (epochs), from a Philox start and from a caller-provided start."""
    n, k, m, lits, off = _instance()
    flat = lits.reshape(-1)
    for devs, name in device_lists():
        with capi.MultiSolver(devs, flags=flags) as ms:
            ms.upload_fixedk(n, lits)
            info = ms.info()
            assert info["sharded"] and info["devices_in_use"] == len(devs), (name, info)
            for seed in (4, 5):
                ms.randomize(seed)
                st = ms.solve(seed)
                v = oracle.randomize(n, seed)
                so = oracle.solve(n, off, flat, v, seed)
                assert st.status == 0 == so.status, name
                assert (st.n_iterations, st.n_resamples, st.sum_mis_size, st.avg_mis_size) == \
                       (so.n_iterations, so.n_resamples, so.sum_mis_size, so.avg_mis_size), name
                assert np.array_equal(ms.get_assignment(), v), name
                assert ms.verify() and oracle.verify(off, flat, v)
                if st.n_incremental_rounds:
                    assert st.n_clause_evals < m * st.n_iterations
                else:
                    assert st.n_clause_evals == m * st.n_iterations
            # caller-provided start (the reference's var_arr->vars) and a round cap
            start = np.random.default_rng(1).integers(0, 2, n, dtype=np.uint8)
            ms.set_assignment(start)
            st = ms.solve(9, max_rounds=3)
            v = start.copy()
            so = oracle.solve(n, off, flat, v, 9, max_rounds=3)
            assert (st.status, st.n_iterations, st.n_resamples) == (so.status, so.n_iterations, so.n_resamples), name
            assert np.array_equal(ms.get_assignment(), v), name
            # every replica holds the same assignment
            for i in range(len(devs)):
                assert np.array_equal(ms.device_solver(i).get_assignment(), v), (name, i)


def test_small_and_unshardable_instances_fall_back_to_the_first_device(capi, oracle, golden):
    """Tiny instances, k > 8 and ragged input are solved on the first device alone -- same results, info() says so;
    ALLL_FLAG_FORCE_SHARDING shards a small uniform instance anyway (one clause per device is enough)."""
    from conftest import golden_case

    n, off, lit, _ = golden_case(golden, "cfg1")              # m ~ 1,200 clauses: below the 4096-per-device threshold
    for flags, want_sharded in ((0, False), (capi.FLAG_FORCE_SHARDING, True)):
        with capi.MultiSolver([0, 0], flags=flags) as ms:
            ms.upload_csr(n, off, lit)                       # uniform width -> routed to the fixed-width upload
            assert ms.info()["sharded"] == want_sharded
            ms.randomize(3)
            st = ms.solve(3)
            v = oracle.randomize(n, 3)
            so = oracle.solve(n, off, lit, v, 3)
            assert (st.status, st.n_iterations, st.n_resamples, st.sum_mis_size) == (0, so.n_iterations, so.n_resamples, so.sum_mis_size)
            assert np.array_equal(ms.get_assignment(), v) and ms.verify()
    n, off, lit, _ = golden_case(golden, "ragged")
    with capi.MultiSolver([0, 0, 0]) as ms:
        ms.upload_csr(n, off, lit)
        assert not ms.info()["sharded"]
        ms.randomize(1)
        st = ms.solve(1, 300)
        v = oracle.randomize(n, 1)
        so = oracle.solve(n, off, lit, v, 1, max_rounds=300)
        assert (st.status, st.n_iterations, st.n_resamples) == (so.status, so.n_iterations, so.n_resamples)
        assert np.array_equal(ms.get_assignment(), v)
    with capi.MultiSolver([0]) as ms:                        # a list of one device is the single-GPU solver
        ms.upload_fixedk(n := 5000, (np.arange(15000, dtype=np.uint32).reshape(5000, 3) % (2 * n)))
        assert ms.info() == dict(devices_in_use=1, sharded=False, widest_range=5000, cap_records=0)
        ms.randomize(0)
        assert ms.solve(0).status == 0 and ms.verify()


def test_exchange_capacity_overflow_is_reported(capi):
    """A start that violates far more clauses than the exchange slots hold ends the solve with ALLL_CAPACITY on every
    rank (abort word), not with a hang or a wrong result."""
    # (k = 5: the random start of the second solve violates m / 32 = 6 k clauses in total -- simulated ranks sharing ONE
    # GPU must stay on the cluster-sized independent-set path, |U| <= 8192: a grid-wide cooperative phase of one rank
    # cannot become resident while the other rank's kernel spins on the same GPU waiting for it)
    n, m, k = 1_000_000, 200_000, 5
    lits = (np.arange(m * k, dtype=np.uint32).reshape(m, k) * 2)          # all-positive, disjoint clauses
    with capi.MultiSolver([0, 0]) as ms:
        ms.upload_fixedk(n, lits)
        assert ms.info()["sharded"] and ms.info()["cap_records"] < m // 2
        ms.set_assignment(np.zeros(n, np.uint8))                          # every clause violated
        with pytest.raises(capi.AlllError) as e:
            ms.solve(0)
        assert e.value.status == capi.CAPACITY
        ms.randomize(1)                                                   # the handle stays usable
        assert ms.solve(1).status == 0 and ms.verify()


def test_absent_peer_times_out_instead_of_hanging(capi):
    """Liveness at world > 1 (round-1 finding): a rank whose peer never starts its solve gives up after the configured
    wait (ALLL_P2P_TIMEOUT_MS) with an error that names the time-out, and raises the abort word for its peers."""
    n, k, m, lits, off = _instance(n=60_000)
    os.environ["ALLL_P2P_TIMEOUT_MS"] = "300"
    try:
        with capi.MultiSolver([0, 0]) as ms:
            ms.upload_fixedk(n, lits)                        # (the time-out is read when the ranks are linked)
            ms.randomize(2)
            lone = ms.device_solver(0)                       # rank 0 solves, rank 1 never shows up
            t0 = time.time()
            with pytest.raises(capi.AlllError) as e:
                lone.solve_p2p(2, m, epoch=77)
            assert e.value.status == capi.CUDA_ERROR and "in time" in str(e.value)
            assert time.time() - t0 < 10.0
    finally:
        del os.environ["ALLL_P2P_TIMEOUT_MS"]


def test_batch_and_portfolio_over_a_device_list(capi, oracle):
    """alll_multi_batch_*: instance blocks per device == alll_batch_solve on one device, job for job; the portfolio over
    all devices has exactly one winner, whose assignment is the plain solve of that seed."""
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat
    from oracle.oracle import to_csr

    n, k, d, n_inst = 2000, 5, 3, 96
    mats = [bounded_degree_ksat(n, k, d, seed=100 + i) for i in range(n_inst)]
    clause_off = np.zeros(n_inst + 1, np.uint64)
    clause_off[1:] = np.cumsum([x.shape[0] for x in mats])
    lits = np.concatenate(mats)
    seeds = np.arange(500, 500 + n_inst, dtype=np.uint64)
    with capi.Solver() as one:
        one.batch_upload(n, k, clause_off, lits)
        want, want_assign, _, _ = one.batch_solve(seeds)
    for devs, name in device_lists():
        with capi.MultiSolver(devs) as ms:
            ms.batch_upload(n, k, clause_off, lits)
            got, assign, _, ms_t = ms.batch_solve(seeds)
            for f in ("n_iterations", "n_resamples", "sum_mis_size", "status"):
                assert np.array_equal(got[f], want[f]), (name, f, np.flatnonzero(got[f] != want[f])[:8])
            assert np.array_equal(assign, want_assign), (name, np.flatnonzero((assign != want_assign).any(axis=1))[:8])
            assert ms_t > 0, (name, ms_t)
            # seed portfolio on instance 0
            ms.batch_upload(n, k, clause_off[:2], lits)
            pseeds = np.arange(9000, 9000 + 300, dtype=np.uint64)
            stats, passign, winner, _ = ms.batch_solve(pseeds, portfolio=True)
            won = np.flatnonzero(stats["status"] == 0)
            assert len(won) == 1 and winner == int(won[0]), (name, won, winner)
            assert set(np.unique(stats["status"]).tolist()) <= {0, capi.PREEMPTED}, (name, np.unique(stats["status"]))
            off0, lit0 = to_csr(mats[0])
            v = oracle.randomize(n, int(pseeds[winner]))
            oracle.solve(n, off0, lit0, v, int(pseeds[winner]))
            assert oracle.verify(off0, lit0, v), name
            assert np.array_equal(passign[winner], v), (name, winner, int((passign[winner] != v).sum()))
