#!/usr/bin/env python
"""Determinism stress of the batched path: the same batch and seeds solved repeatedly must give identical statistics and
assignments every time (the kernels resolve their races with atomics; any order dependence would show up here)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alllsatisfiabilitysolver_b200 import capi
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=1024)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
out = []
for (n, k, d) in [(10_000, 5, 3), (3000, 7, 20), (1500, 3, 4), (9000, 3, 3), (4000, 8, 30), (3000, 4, 3)]:
    n_inst = a.instances if n * d // k <= 6000 else max(32, a.instances // 8)
    insts = [bounded_degree_ksat(n, k, d, seed=0xB000 + i) for i in range(n_inst)]
    off = np.zeros(n_inst + 1, np.uint64); off[1:] = np.cumsum([x.shape[0] for x in insts])
    lits = np.concatenate(insts, axis=0)
    seeds = np.arange(n_inst, dtype=np.uint64) + np.uint64(31)
    with capi.Solver(device=0) as s:
        s.batch_upload(n, k, off, lits)
        ref = None
        for r in range(a.reps):
            stats, assign, _, ms = s.batch_solve(seeds)
            cur = (stats["n_iterations"].copy(), stats["n_resamples"].copy(), stats["sum_mis_size"].copy(), stats["status"].copy(), assign.copy())
            if ref is None:
                ref = cur
            else:
                for x, y in zip(ref, cur):
                    assert np.array_equal(x, y), f"run {r} differs from run 0 at shape {(n, k, d)}"
        assert (ref[3] == 0).all()
    out.append(dict(n=n, k=k, d=d, instances=n_inst, reps=a.reps, identical=True, last_ms=ms))
print(json.dumps(out))
