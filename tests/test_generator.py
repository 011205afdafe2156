"""Enumerated-clause solve (SATInstance.h:70-153, ClauseGenerator.h:16-114): the clauses are a pure
function of their index and are never stored.  Parity: the oracle solves the MATERIALISED instance (clauses written out
by its own restatement of the generator); violated sets, independent sets, statistics and the final assignment of the
device path must be identical."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alllsatisfiabilitysolver_b200")

#        kind  n_vars    m      k  seed  d
CASES = [
    (1, 2_000, 1_200, 5, 11, 3),          # cfg1 shape, bounded occurrence
    (0, 5_000, 9_000, 6, 12, 0),          # uniform 6-SAT, ratio 1.8
    (1, 50_000, 190_000, 8, 13, 32),      # cfg4 shape scaled down: |U| of round 0 stays within one cluster
    (1, 400_000, 1_600_000, 8, 14, 32),   # round 0 goes through the cooperative grid path
    (0, 3_000, 4_000, 3, 15, 0),          # narrow clauses, many rounds
    (1, 1_000, 900, 16, 16, 15),          # widest built-in
]


@pytest.fixture(scope="module")
def oracle():
    return orc.Oracle()


def test_library_host_generators_match_the_oracle_restatement(oracle):
    """alll_builtin_generator_clause (the functor the device kernels run, compiled for the host) against
    oracle/alll_oracle.c:alll_oracle_gen_materialize (plain % arithmetic written from the header's specification)."""
    from alllsatisfiabilitysolver_b200 import capi
    rng = np.random.default_rng(3)
    for kind, n, m, k, seed, d in CASES + [(1, 10_000_000, 40_000_000, 8, 17, 32)]:
        idx = np.arange(m) if m <= 10_000 else np.unique(np.concatenate([rng.integers(0, m, 3000), [0, m - 1]]))
        got = capi.builtin_generator_clauses(kind, n, m, k, seed, d, indices=idx)
        if m <= 2_000_000:
            exp = oracle.gen_materialize(kind, n, m, k, seed, d)
            assert (got == exp[idx]).all()
            if kind == 1:
                assert np.bincount((exp >> 1).ravel()).max() <= d
        assert (got >> 1).max() < n


def test_generator_parameter_checks():
    from alllsatisfiabilitysolver_b200 import capi
    for bad in [(1, 100, 100, 5, 1, 3), (1, 100, 10, 17, 1, 3), (2, 100, 10, 3, 1, 3), (1, 100, 10, 3, 1, 0)]:
        with pytest.raises(capi.AlllError):
            capi.builtin_generator_clauses(*bad, indices=[0])


def test_user_translation_unit_compiles_for_sm100a(tmp_path):
    """include/alll_generator.cuh is a public header: a user TU must build with nothing but include/ on the path."""
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "csrc")], check=True)
    out = str(tmp_path / "user_generator")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "user_generator.cu"), "-o", out, "-L" + PKG, "-lalll_b200",
                    "-Xlinker", "-rpath," + PKG], check=True)
    assert os.path.exists(out)


# ---- GPU ---------------------------------------------------------------------------------------------

@pytest.fixture(scope="module")
def capi():
    from alllsatisfiabilitysolver_b200 import capi as c
    return c


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[f"kind{c[0]}-n{c[1]}-k{c[3]}" for c in CASES])
def test_enumerated_solve_equals_oracle_on_the_materialised_instance(capi, oracle, case):
    kind, n, m, k, seed, d = case
    lits = oracle.gen_materialize(kind, n, m, k, seed, d)
    off, lit = orc.to_csr(lits)
    s = capi.Solver()
    s.upload_builtin_generator(kind, n, m, k, seed, d)
    # violated set of a fixed assignment (Clause.h:34-46 on every enumerated clause)
    vars0 = oracle.randomize(n, seed + 1)
    s.set_assignment(vars0)
    n_viol, ids = s.eval()
    exp_u = oracle.sweep(off, lit, vars0)
    assert n_viol == len(exp_u) and sorted(ids.tolist()) == exp_u.tolist()
    # one round: U, S and the resampled assignment
    got_u, got_s, got_r = s.round(seed, 0)
    work = vars0.copy()
    u, sset, nres = oracle.round(n, off, lit, work, seed, 0)
    assert sorted(got_u.tolist()) == sorted(u.tolist())
    assert sorted(got_s.tolist()) == sorted(sset.tolist())
    assert got_r == nres
    assert (s.get_assignment() == work).all()
    # whole solve from the seeded assignment
    s.randomize(seed)
    st = s.solve(seed, max_rounds=3000)
    vars1 = oracle.randomize(n, seed)
    exp = oracle.solve(n, off, lit, vars1, seed, max_rounds=3000)
    assert (st.status, st.n_iterations, st.n_resamples, st.avg_mis_size) == (exp.status, exp.n_iterations, exp.n_resamples, exp.avg_mis_size)
    got = s.get_assignment()
    assert (got == vars1).all()
    if st.status == 0:
        assert oracle.verify(off, lit, got) and s.verify()
    s.close()


@pytest.mark.gpu
def test_enumerated_solve_equals_stored_solve(capi, oracle):
    """The same clauses uploaded as a literal matrix and as a generator: identical statistics and assignment."""
    kind, n, m, k, seed, d = 1, 300_000, 1_200_000, 8, 21, 32
    lits = oracle.gen_materialize(kind, n, m, k, seed, d)
    a, b = capi.Solver(), capi.Solver()
    a.upload_fixedk(n, lits)
    b.upload_builtin_generator(kind, n, m, k, seed, d)
    for s in (a, b):
        s.randomize(5)
    sa, sb = a.solve(5), b.solve(5)
    assert sa.status == sb.status == 0
    assert (sa.n_iterations, sa.n_resamples, sa.sum_mis_size) == (sb.n_iterations, sb.n_resamples, sb.sum_mis_size)
    assert (a.get_assignment() == b.get_assignment()).all()
    a.close(); b.close()


@pytest.mark.gpu
def test_record_capacity_overflow_is_reported(capi):
    s = capi.Solver()
    s.upload_builtin_generator(1, 50_000, 190_000, 8, 13, 32, cap_records=16)     # round 0 violates ~740 clauses
    s.randomize(1)
    with pytest.raises(capi.AlllError) as e:
        s.solve(1)
    assert "CAPACITY" in str(e.value)
    with pytest.raises(capi.AlllError):
        s.randomize(1)
        s.eval()
    # a sufficient capacity on the same handle works again
    s.upload_builtin_generator(1, 50_000, 190_000, 8, 13, 32, cap_records=4096)
    s.randomize(1)
    assert s.solve(1).status == 0 and s.verify()
    s.close()


@pytest.mark.gpu
def test_user_generator_program(oracle, tmp_path):
    """tests/cpp/user_generator.cu: a user functor compiled into the public sweep kernel, solved through the C ABI;
    the oracle solves the same clauses written out with numpy."""
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "csrc")], check=True)
    exe = str(tmp_path / "user_generator")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "user_generator.cu"), "-o", exe, "-L" + PKG, "-lalll_b200",
                    "-Xlinker", "-rpath," + PKG], check=True)
    n, m, seed, k = 30_000, 14_000, 77, 5
    out = str(tmp_path / "vars.bin")
    got = json.loads(subprocess.run([exe, str(n), str(m), str(seed), out], capture_output=True, text=True, check=True).stdout)
    i = np.arange(m, dtype=np.uint64)
    h = (i.astype(np.uint32) * np.uint32(2654435761)).astype(np.uint32)
    lits = np.stack([2 * ((i + j * 7919) % n).astype(np.uint32) + ((h >> np.uint32(j + 7)) & 1) for j in range(k)], axis=1).astype(np.uint32)
    off, lit = orc.to_csr(lits)
    vars1 = oracle.randomize(n, seed)
    exp = oracle.solve(n, off, lit, vars1, seed, max_rounds=10000)
    assert got["valid"] == 1 and got["host_violated"] == 0
    assert (got["n_iterations"], got["n_resamples"], got["avg_mis_size"]) == (exp.n_iterations, exp.n_resamples, exp.avg_mis_size)
    assert (np.fromfile(out, np.uint8) == vars1).all()
