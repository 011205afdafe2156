"""Live cross-check of the C restatement against the reference shim on fresh random inputs.
Runs wherever oracle/_ref/liballl_ref.so exists (built here; prebuilt on the GPU box). CPU only."""
import numpy as np
import pytest

from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat, uniform_ksat
from oracle.oracle import to_csr


@pytest.mark.parametrize("shape", [("b", 4000, 5, 3), ("b", 3000, 7, 28), ("u", 2000, 3, 5000), ("b", 6000, 8, 32)])
@pytest.mark.parametrize("nt", [1, 4])
def test_sweep_and_greedy_live(oracle, reference, shape, nt):
    kind, n, k, x = shape
    lits = bounded_degree_ksat(n, k, x, seed=99) if kind == "b" else uniform_ksat(n, k, x, seed=99)
    off, lit = to_csr(lits)
    rng = np.random.default_rng(1)
    with reference.instance(n, off, lit, nt) as ri:
        for _ in range(3):
            v = rng.integers(0, 2, n, dtype=np.uint8)
            ri.set_assignment(v)
            u = ri.sweep()
            assert np.array_equal(u, oracle.sweep(off, lit, v))
            assert ri.verify() == oracle.verify(off, lit, v)
            assert np.array_equal(ri.greedy_mis(), oracle.greedy_mis(off, lit, u, nt))


def test_oracle_solution_passes_reference_verify(oracle, reference):
    n, lits = 2000, bounded_degree_ksat(2000, 5, 3, seed=5)
    off, lit = to_csr(lits)
    v = oracle.randomize(n, 42)
    st = oracle.solve(n, off, lit, v, 42)
    assert st.status == 0
    with reference.instance(n, off, lit, 2) as ri:
        ri.set_assignment(v)
        assert ri.verify() and len(ri.sweep()) == 0


def test_reference_solve_runs(reference):
    n, lits = 2000, bounded_degree_ksat(2000, 5, 3, seed=5)
    off, lit = to_csr(lits)
    with reference.instance(n, off, lit, 2) as ri:
        st = ri.solve()
        assert ri.verify() and st.n_iterations >= 1
