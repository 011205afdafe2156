"""Pins oracle/alll_oracle.c against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py -> oracle/_ref).  CPU only."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_case

CASES = ["cfg1", "k7_small", "k8_small", "k3_uniform", "ragged", "tiny"]

# Random123 v1.14 known-answer vectors for Philox4x32-10 (kat_vectors: counter, key -> output)
PHILOX_KAT = [
    ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
    ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
    ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
     [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
]


def test_philox_known_answers(oracle):
    for ctr, key, want in PHILOX_KAT:
        assert [int(x) for x in oracle.philox(ctr, key)] == want


@pytest.mark.parametrize("name", CASES)
def test_sweep_matches_reference(oracle, golden, name):
    """Clause.h:34-46 / SATInstance.h:273-280: violated sets bit-exact."""
    n, off, lit, assigns = golden_case(golden, name)
    for a in range(assigns.shape[0]):
        u = oracle.sweep(off, lit, assigns[a])
        assert np.array_equal(u, golden[f"{name}/U{a}"])
        assert oracle.verify(off, lit, assigns[a]) == bool(golden[f"{name}/valid{a}"][0])


@pytest.mark.parametrize("name", CASES)
def test_dependent_matches_reference(oracle, golden, name):
    """SATInstance.h:369-389"""
    n, off, lit, _ = golden_case(golden, name)
    pairs, dep = golden[f"{name}/dep_pairs"], golden[f"{name}/dep"]
    for (a, b), d in zip(pairs, dep):
        la = lit[int(off[a]):int(off[a + 1])]
        lb = lit[int(off[b]):int(off[b + 1])]
        assert oracle.dependent(la, lb) == bool(d)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("nt", [1, 3, 8])
def test_greedy_mis_matches_reference(oracle, golden, name, nt):
    """SATInstance.h:391-451 incl. the round-robin start at list 1 and the batch split of main.cpp:173."""
    n, off, lit, assigns = golden_case(golden, name)
    for a in range(assigns.shape[0]):
        key = f"{name}/greedy{a}_t{nt}"
        if key not in golden:
            continue
        got = oracle.greedy_mis(off, lit, golden[f"{name}/U{a}"], nt)
        assert np.array_equal(got, golden[key])


def test_check_signed_matches_cnf_evaluate(oracle, golden):
    """cnf_io.cpp:392-484 on the raw signed literals the reference parser produced."""
    exp = json.load(open(os.path.join(GOLDEN, "dimacs", "expected.json")))["cfg1.cnf"]
    assigns = golden["cfg1/assign"]
    for a, want in enumerate(exp["cnf_evaluate_assign"]):
        assert oracle.check_signed(exp["l_c_num"], exp["l_val"], assigns[a]) == want


def test_priority_mis_is_independent_and_maximal(oracle, golden):
    """The two properties the reference's greedy result has by construction (SATInstance.h:415-447)."""
    for name in CASES:
        n, off, lit, assigns = golden_case(golden, name)
        for a in range(assigns.shape[0]):
            u = golden[f"{name}/U{a}"]
            s = oracle.priority_mis(n, off, lit, u, seed=5, rnd=a)
            used = {}
            for c in s:
                for l in lit[int(off[c]):int(off[c + 1])]:
                    assert used.setdefault(int(l) >> 1, int(c)) == int(c)       # independence
            in_s = set(int(c) for c in s)
            for c in u:
                if int(c) in in_s:
                    continue
                vs = [int(l) >> 1 for l in lit[int(off[c]):int(off[c + 1])]]
                assert any(v in used for v in vs) or len(vs) == 0 and False       # maximality


def test_solve_statistics_semantics(oracle, golden):
    """n_iterations = rounds+1, n_resamples counts variables, avg_mis_size = floor(sum|S|/n_iterations)
    (SATInstance.h:261,285-287,291,317,363)."""
    n, off, lit, _ = golden_case(golden, "cfg1")
    v = oracle.randomize(n, 3)
    st, tu, ts = oracle.solve(n, off, lit, v, 3, trace=True)
    assert st.status == 0 and oracle.verify(off, lit, v)
    assert st.n_iterations == len(tu) and tu[-1] == 0 and (tu[:-1] > 0).all()
    assert st.sum_mis_size == int(ts.sum())
    assert st.avg_mis_size == int(ts.sum()) // st.n_iterations
    assert st.n_resamples == 5 * int(ts.sum())          # k = 5 everywhere in cfg1
    assert st.n_clause_evals == (len(off) - 1) * st.n_iterations


def test_round_counts_match_reference_distribution(oracle, golden):
    """north_star: resample-round counts match the reference's parallel MT distribution within a stated
    tolerance.  Tolerance (calibrated on the golden reference runs, n=300): mean n_iterations within
    +-15 %, mean n_resamples within +-5 %."""
    n, off, lit, _ = golden_case(golden, "cfg1")
    ref_it = golden["cfg1/ref_solve_t1_iterations"].astype(float)
    ref_rs = golden["cfg1/ref_solve_t1_resamples"].astype(float)
    its, rs = [], []
    for seed in range(300):
        v = oracle.randomize(n, seed)
        st = oracle.solve(n, off, lit, v, seed)
        assert st.status == 0
        its.append(st.n_iterations)
        rs.append(st.n_resamples)
    assert abs(np.mean(its) - ref_it.mean()) <= 0.15 * ref_it.mean()
    assert abs(np.mean(rs) - ref_rs.mean()) <= 0.05 * ref_rs.mean()
    # and the greedy-MIS restatement of the reference's own law
    its_g = []
    for seed in range(300):
        v = oracle.randomize(n, 1000 + seed)
        st = oracle.solve_greedy(n, off, lit, v, seed)
        its_g.append(st.n_iterations)
    assert abs(np.mean(its_g) - ref_it.mean()) <= 0.10 * ref_it.mean()


def test_empty_clause_and_round_cap(oracle):
    off = np.array([0, 1, 1], np.uint64)
    lit = np.array([0], np.uint32)
    v = np.zeros(1, np.uint8)
    assert oracle.sweep(off, lit, v).tolist() == [0, 1]           # empty clause is always violated (Clause.h:35-45)
    assert oracle.solve(1, off, lit, v, 0).status == 2            # EMPTY_CLAUSE instead of the reference's livelock
    # x and not-x: unsatisfiable -> round cap
    off = np.array([0, 1, 2], np.uint64)
    lit = np.array([0, 1], np.uint32)
    st = oracle.solve(1, off, lit, np.zeros(1, np.uint8), 0, max_rounds=50)
    assert st.status == 1 and st.n_iterations == 50
