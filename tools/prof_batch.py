#!/usr/bin/env python
"""BASELINE config 5: N x 5-SAT n=10k (d=3) solved by the one-CTA-per-instance kernel; prints instances/s."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alllsatisfiabilitysolver_b200 import capi
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat

ap = argparse.ArgumentParser()
ap.add_argument("--instances", type=int, default=8192)
ap.add_argument("--n", type=int, default=10_000)
ap.add_argument("--k", type=int, default=5)
ap.add_argument("--d", type=int, default=3)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--portfolio", action="store_true")
a = ap.parse_args()
t = time.time()
insts = [bounded_degree_ksat(a.n, a.k, a.d, seed=0xA116 + i) for i in range(1 if a.portfolio else a.instances)]
off = np.zeros(len(insts) + 1, np.uint64); off[1:] = np.cumsum([x.shape[0] for x in insts])
lits = np.concatenate(insts, axis=0)
gen_s = time.time() - t
s = capi.Solver(device=0)
t = time.time(); s.batch_upload(a.n, a.k, off, lits); up_s = time.time() - t
out = dict(instances=a.instances, n=a.n, k=a.k, total_clauses=int(lits.shape[0]), gen_s=gen_s, upload_s=up_s, runs=[])
for r in range(a.reps):
    seeds = np.arange(r * a.instances, (r + 1) * a.instances, dtype=np.uint64)
    stats, assign, winner, ms = s.batch_solve(seeds, portfolio=a.portfolio, want_assignments=(r == 0))
    ok = int((stats["status"] == 0).sum())
    out["runs"].append(dict(ms=ms, solved=ok, winner=winner, mean_iters=float(stats["n_iterations"].mean()),
                            instances_per_s=a.instances / (ms * 1e-3),
                            clause_evals_per_s=float((stats["n_iterations"] * (lits.shape[0] / len(insts))).sum()) / (ms * 1e-3)))
print(json.dumps(out))
