"""Generates tests/golden/*.npz and tests/golden/dimacs/* from the UNMODIFIED reference.

Run in the build container (where /root/reference exists):

    make -C oracle && python tests/golden/make_golden.py

Every expected value below comes from oracle/_ref/liballl_ref.so, i.e. from the
reference's own Clause::is_not_satisfied (Clause.h:34), dependent_clauses
(SATInstance.h:369), populate_mis_parallel (SATInstance.h:391), SATInstance::solve
(SATInstance.h:60), verify_validity (SATInstance.h:156) and cnf_io
(cnf_io.cpp:126,487,392) -- never from this repository's own code.  The inputs
(instances, assignments) are seeded numpy draws stored alongside the outputs, so
the fixtures are self-contained and do not need the reference at test time.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from alllsatisfiabilitysolver_b200.instances import (  # noqa: E402
    bounded_degree_ksat, make_config, uniform_ksat, write_dimacs)
from oracle.oracle import Reference, to_csr  # noqa: E402


def ragged_instance(rng, n, m):
    """Variable-width clauses incl. unit clauses, duplicated literals and tautologies."""
    clauses = []
    for _ in range(m):
        k = int(rng.integers(1, 9))
        vs = rng.integers(0, n, size=k)
        sg = rng.integers(0, 2, size=k)
        c = list((vs * 2 + sg).astype(np.uint32))
        r = rng.random()
        if r < 0.1 and k >= 2:
            c[1] = c[0]              # duplicated literal
        elif r < 0.2 and k >= 2:
            c[1] = c[0] ^ 1          # x or not x
        clauses.append(c)
    return clauses


def main():
    R = Reference()
    rng = np.random.default_rng(20261018)
    out = {}
    cases = []

    def add_case(name, n, lits, n_assign=4):
        off, lit = to_csr(lits)
        out[f"{name}/n_vars"] = np.array([n], np.uint64)
        out[f"{name}/off"] = off
        out[f"{name}/lit"] = lit
        m = len(off) - 1
        assigns = rng.integers(0, 2, size=(n_assign, n), dtype=np.uint8)
        # one all-false and one all-true assignment as edge cases
        assigns[0, :] = 0
        if n_assign > 1:
            assigns[1, :] = 1
        out[f"{name}/assign"] = assigns
        for nt in (1, 3, 8):
            with R.instance(n, off, lit, nt) as ri:
                for a in range(n_assign):
                    ri.set_assignment(assigns[a])
                    u = ri.sweep()
                    if nt == 1:
                        out[f"{name}/U{a}"] = u
                        out[f"{name}/valid{a}"] = np.array([ri.verify()], np.uint8)
                    else:
                        assert np.array_equal(u, out[f"{name}/U{a}"])
                    if len(u) <= 20000:
                        out[f"{name}/greedy{a}_t{nt}"] = ri.greedy_mis()
                if nt == 1 and m >= 2:
                    pairs = rng.integers(0, m, size=(256, 2)).astype(np.uint32)
                    dep = np.array([ri.dependent(int(a), int(b)) for a, b in pairs], np.uint8)
                    out[f"{name}/dep_pairs"] = pairs
                    out[f"{name}/dep"] = dep
        cases.append(name)

    n1, l1 = make_config("cfg1")
    add_case("cfg1", n1, l1)
    add_case("k7_small", 5000, bounded_degree_ksat(5000, 7, 28, 11))
    add_case("k8_small", 20000, bounded_degree_ksat(20000, 8, 32, 12))
    add_case("k3_uniform", 3000, uniform_ksat(3000, 3, 9000, 13))
    add_case("ragged", 300, ragged_instance(rng, 300, 700))
    add_case("tiny", 3, [[0], [1, 2], [5, 5, 4]], n_assign=3)

    # ---- reference solve statistics (self-seeded; distributional calibration) ----
    off, lit = to_csr(l1)
    for nt, runs in ((1, 300), (8, 100)):
        its, res, mis = [], [], []
        with R.instance(n1, off, lit, nt) as ri:
            for _ in range(runs):
                ri.rerandomize()
                st = ri.solve()
                assert ri.verify()
                its.append(st.n_iterations)
                res.append(st.n_resamples)
                mis.append(st.avg_mis_size)
        out[f"cfg1/ref_solve_t{nt}_iterations"] = np.array(its, np.uint32)
        out[f"cfg1/ref_solve_t{nt}_resamples"] = np.array(res, np.uint32)
        out[f"cfg1/ref_solve_t{nt}_avg_mis"] = np.array(mis, np.uint32)
    n7, l7 = 100_000, bounded_degree_ksat(100_000, 7, 28, 17)
    off7, lit7 = to_csr(l7)
    its, res = [], []
    with R.instance(n7, off7, lit7, 8) as ri:
        for _ in range(30):
            ri.rerandomize()
            st = ri.solve()
            assert ri.verify()
            its.append(st.n_iterations)
            res.append(st.n_resamples)
    out["k7_100k/ref_solve_t8_iterations"] = np.array(its, np.uint32)
    out["k7_100k/ref_solve_t8_resamples"] = np.array(res, np.uint32)
    out["k7_100k/shape"] = np.array([n7, 7, 28, 17, l7.shape[0]], np.uint64)  # n,k,d,seed,m

    out["cases"] = np.array(cases)
    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)

    # ---- DIMACS dialect fixtures parsed by the reference's cnf_io ----
    ddir = os.path.join(HERE, "dimacs")
    os.makedirs(ddir, exist_ok=True)
    files = {}
    write_dimacs(os.path.join(ddir, "cfg1.cnf"), n1, l1, comment="cfg1: 5-SAT n=2000 d=3 seed 0xA112")
    files["cfg1.cnf"] = None
    with open(os.path.join(ddir, "dialect.cnf"), "w") as f:
        f.write("c leading comment\nC upper-case comment\n\np  CNF   6 5\n1 -2 3 0\nc comment between clauses\n"
                "4 5\n-6 0 -1 2 0\n   3   -4   0\n6 -5\n 1 0\n")
    files["dialect.cnf"] = None
    with open(os.path.join(ddir, "no_trailing_newline.cnf"), "w") as f:
        f.write("p cnf 4 3\n1 2 0\n-3 4 0\n-1 -4 0")      # last line dropped by the reference
    files["no_trailing_newline.cnf"] = None
    expected = {}
    for name in files:
        v, c, l, l_c_num, l_val = R.cnf_read(os.path.join(ddir, name))
        expected[name] = dict(v_num=v, c_num=c, l_num=l,
                              l_c_num=[int(x) for x in l_c_num[:c]],
                              l_val=[int(x) for x in l_val[:l]])
    # cnf_evaluate on cfg1 for a few assignments
    v, c, l, l_c_num, l_val = R.cnf_read(os.path.join(ddir, "cfg1.cnf"))
    ev = []
    a = np.load(os.path.join(HERE, "golden_v1.npz"))["cfg1/assign"]
    for i in range(a.shape[0]):
        ev.append(bool(R.cnf_evaluate(v, l_c_num[:c], l_val[:l], a[i])))
    expected["cfg1.cnf"]["cnf_evaluate_assign"] = ev
    with open(os.path.join(ddir, "expected.json"), "w") as f:
        json.dump(expected, f)
    print("wrote", os.path.join(HERE, "golden_v1.npz"), "and", ddir)


if __name__ == "__main__":
    main()
