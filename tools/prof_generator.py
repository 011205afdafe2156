"""Enumerated-clause mode on cfg4's shape (8-SAT, n=10M, every variable <= 32 occurrences, m=40M): nothing stored.

    python tools/prof_generator.py [--kind 1] [--n 10000000] [--k 8] [--d 32] [--solves 3]
"""
import argparse, json, os, sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from alllsatisfiabilitysolver_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--kind", type=int, default=1)
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--k", type=int, default=8)
ap.add_argument("--d", type=int, default=32)
ap.add_argument("--m", type=int, default=0)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--solves", type=int, default=3)
ap.add_argument("--cap", type=int, default=0)
args = ap.parse_args()
m = args.m or args.n * args.d // args.k
s = capi.Solver()
s.upload_builtin_generator(args.kind, args.n, m, args.k, 0xA111, args.d, cap_records=args.cap or max(4096, 4 * m >> args.k))
s.randomize(1)
ms, n_viol = s.time_sweep(args.reps)
res = {"kind": args.kind, "n": args.n, "m": m, "k": args.k, "sweep_ms": ms, "clause_evals_per_s": m / ms * 1e3, "n_violated": n_viol, "solves": []}
for seed in range(args.solves):
    s.randomize(seed)
    st = s.solve(seed, max_rounds=2000)
    res["solves"].append({"seed": seed, "status": st.status, "n_iterations": st.n_iterations, "solve_ms": st.solve_ms,
                          "sweep_ms": st.sweep_ms, "between_ms": st.between_sweeps_ms, "verified": bool(s.verify())})
print(json.dumps(res))
