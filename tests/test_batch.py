"""Batched small instances and the seed portfolio (BASELINE config 5)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    from alllsatisfiabilitysolver_b200 import capi as m

    m.load()
    return m


def make_batch(n_inst, n, k, d, seed0):
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

    insts = [bounded_degree_ksat(n, k, d, seed=seed0 + i) for i in range(n_inst)]
    off = np.zeros(n_inst + 1, np.uint64)
    off[1:] = np.cumsum([x.shape[0] for x in insts])
    return insts, off, np.concatenate(insts, axis=0)


@pytest.mark.parametrize("shape", [(48, 2000, 5, 3), (24, 10_000, 5, 3), (16, 3000, 7, 20), (8, 1500, 3, 4),
                                   (6, 9000, 3, 3),       # |U| ~ 1,100 in round 0: every job outgrows the small kernel's records
                                   (4, 2000, 9, 40),      # k = 9: large kernel only
                                   (5, 4000, 8, 30), (5, 3000, 4, 3), (5, 3000, 6, 8)])
def test_batch_matches_oracle_per_instance(capi, oracle, shape):
    """Every instance of the batch ends with the oracle's assignment and Statistics for its seed
    (same round specification as alll_solve), and the assignment satisfies the instance."""
    n_inst, n, k, d = shape
    insts, off, lits = make_batch(n_inst, n, k, d, 500)
    seeds = np.arange(1000, 1000 + n_inst, dtype=np.uint64)
    with capi.Solver() as s:
        s.batch_upload(n, k, off, lits)
        stats, assign, winner, ms = s.batch_solve(seeds)
    assert winner == -1 and ms > 0
    for i in range(n_inst):
        m = insts[i].shape[0]
        coff = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
        flat = insts[i].reshape(-1)
        v = oracle.randomize(n, int(seeds[i]))
        so = oracle.solve(n, coff, flat, v, int(seeds[i]))
        assert stats["status"][i] == 0 and so.status == 0
        assert (int(stats["n_iterations"][i]), int(stats["n_resamples"][i]), int(stats["sum_mis_size"][i])) == \
               (so.n_iterations, so.n_resamples, so.sum_mis_size)
        assert np.array_equal(assign[i], v) and oracle.verify(coff, flat, assign[i])


def test_batch_mixed_small_and_retried_jobs(capi, oracle):
    """One batch whose instances differ in size: the small ones finish in the small kernel, the ones whose violated set
    outgrows its records are redone by the large kernel behind it -- every job must still equal the oracle."""
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

    n, k = 9000, 3
    insts = [bounded_degree_ksat(n, k, d, seed=4000 + i) for i, d in enumerate([1, 3, 1, 2, 3, 1, 3])]
    sizes = [x.shape[0] for x in insts]
    assert min(sizes) * 2 ** -k < 400 and max(sizes) * 2 ** -k > 700      # both sides of the 512-record limit
    off = np.zeros(len(insts) + 1, np.uint64)
    off[1:] = np.cumsum(sizes)
    seeds = np.arange(77, 77 + len(insts), dtype=np.uint64)
    with capi.Solver() as s:
        s.batch_upload(n, k, off, np.concatenate(insts, axis=0))
        for _ in range(2):                                                   # the retry list is reset per call
            stats, assign, _, _ = s.batch_solve(seeds)
            for i, lits in enumerate(insts):
                coff = np.arange(lits.shape[0] + 1, dtype=np.uint64) * np.uint64(k)
                v = oracle.randomize(n, int(seeds[i]))
                so = oracle.solve(n, coff, lits.reshape(-1), v, int(seeds[i]))
                assert stats["status"][i] == 0 and so.status == 0
                assert (int(stats["n_iterations"][i]), int(stats["n_resamples"][i]), int(stats["sum_mis_size"][i])) == \
                       (so.n_iterations, so.n_resamples, so.sum_mis_size)
                assert np.array_equal(assign[i], v)


def test_batch_edge_cases_small_kernel(capi, oracle):
    """Ragged instance sizes (0, 1, not a multiple of 4 clauses), repeated variables and tautologies inside a clause,
    variables nobody uses, an odd variable count: every job equals the oracle (k = 3 and k = 8: the 256-thread kernel)."""
    rng = np.random.default_rng(99)
    for k, n in ((3, 41), (8, 77)):
        sizes = [0, 1, 3, 5, 4, 7, 64, 129, 2, 0, 33]
        insts = []
        for m in sizes:
            v = rng.integers(0, n - 5, size=(m, k), dtype=np.int64)          # the last 5 variables never occur
            if m:
                v[0, 1] = v[0, 0]                                            # a repeated variable ...
            lits = (v * 2 + rng.integers(0, 2, size=(m, k), dtype=np.int64)).astype(np.uint32)
            if m > 2:
                lits[2, 1] = lits[2, 0] ^ np.uint32(1)                       # ... and a tautology (x or not-x)
            insts.append(lits)
        off = np.zeros(len(insts) + 1, np.uint64)
        off[1:] = np.cumsum(sizes)
        seeds = np.arange(5, 5 + len(insts), dtype=np.uint64)
        with capi.Solver() as s:
            s.batch_upload(n, k, off, np.concatenate(insts, axis=0))
            stats, assign, _, _ = s.batch_solve(seeds, max_rounds=500)
        for i, lits in enumerate(insts):
            coff = np.arange(lits.shape[0] + 1, dtype=np.uint64) * np.uint64(k)
            vv = oracle.randomize(n, int(seeds[i]))
            so = oracle.solve(n, coff, lits.reshape(-1), vv, int(seeds[i]), 500)
            assert int(stats["status"][i]) == so.status, (k, i)
            assert (int(stats["n_iterations"][i]), int(stats["n_resamples"][i]), int(stats["sum_mis_size"][i])) == \
                   (so.n_iterations, so.n_resamples, so.sum_mis_size), (k, i)
            assert np.array_equal(assign[i], vv), (k, i)


def test_batch_agrees_with_single_instance_path(capi):
    """The one-CTA kernel and the large-instance kernels implement one specification."""
    insts, off, lits = make_batch(6, 4000, 5, 3, 900)
    seeds = np.arange(7, 13, dtype=np.uint64)
    with capi.Solver() as s:
        s.batch_upload(4000, 5, off, lits)
        stats, assign, _, _ = s.batch_solve(seeds)
        for i in range(6):
            s.upload_fixedk(4000, insts[i])
            s.randomize(int(seeds[i]))
            st = s.solve(int(seeds[i]))
            assert (st.n_iterations, st.n_resamples, st.sum_mis_size) == \
                   (int(stats["n_iterations"][i]), int(stats["n_resamples"][i]), int(stats["sum_mis_size"][i]))
            assert np.array_equal(s.get_assignment(), assign[i])


def test_portfolio_first_sat_wins(capi, oracle):
    """One instance, many seeds: exactly one winner; its assignment is the oracle's for that seed and satisfies the
    instance; every other job either was pre-empted or also finished (status OK is impossible for a non-winner)."""
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

    n, k = 10_000, 5
    lits = bounded_degree_ksat(n, k, 3, seed=77)
    m = lits.shape[0]
    coff = np.arange(m + 1, dtype=np.uint64) * np.uint64(k)
    seeds = np.arange(40, 40 + 512, dtype=np.uint64)
    with capi.Solver() as s:
        s.batch_upload(n, k, np.array([0, m], np.uint64), lits)
        stats, assign, winner, ms = s.batch_solve(seeds, portfolio=True)
    assert 0 <= winner < len(seeds)
    assert stats["status"][winner] == capi.OK
    others = np.delete(stats["status"], winner)
    assert set(others.tolist()) <= {capi.PREEMPTED}
    v = oracle.randomize(n, int(seeds[winner]))
    so = oracle.solve(n, coff, lits.reshape(-1), v, int(seeds[winner]))
    assert np.array_equal(assign[winner], v) and oracle.verify(coff, lits.reshape(-1), assign[winner])
    assert int(stats["n_iterations"][winner]) == so.n_iterations


def test_batch_round_cap_and_limits(capi):
    with capi.Solver() as s:
        # x and not-x in every instance: never satisfiable -> MAX_ROUNDS with exactly max_rounds sweeps
        lits = np.tile(np.array([[0], [1]], np.uint32), (4, 1))
        s.batch_upload(1, 1, np.array([0, 2, 4, 6, 8], np.uint64), lits)
        stats, _, _, _ = s.batch_solve(np.arange(4, dtype=np.uint64), max_rounds=25)
        assert (stats["status"] == capi.MAX_ROUNDS).all() and (stats["n_iterations"] == 25).all()
        # too many variables for one CTA's shared memory
        with pytest.raises(capi.AlllError) as e:
            s.batch_upload(1_000_000, 3, np.array([0, 1], np.uint64), np.array([[0, 2, 4]], np.uint32))
        assert e.value.status == capi.BAD_ARG
        # job count must match in non-portfolio mode
        s.batch_upload(10, 2, np.array([0, 1], np.uint64), np.array([[0, 3]], np.uint32))
        with pytest.raises(capi.AlllError):
            s.batch_solve(np.arange(3, dtype=np.uint64))
