#!/usr/bin/env python
"""Short driver for ncu / timing experiments: upload a workload, run `reps` sweeps and a few full solves.

    python tools/prof_sweep.py --workload cfg4 --reps 5 [--solves 1] [--smem BYTES] [--flags F]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from alllsatisfiabilitysolver_b200 import capi  # noqa: E402
from alllsatisfiabilitysolver_b200.instances import CONFIGS, bounded_degree_ksat_torch, uniform_ksat_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--solves", type=int, default=0)
ap.add_argument("--smem", type=int, default=0)
ap.add_argument("--flags", type=int, default=0)
ap.add_argument("--max-rounds", type=int, default=2000)
a = ap.parse_args()
cfg = CONFIGS[a.workload]
n = int(cfg["n"] * a.scale)
if cfg["kind"] == "bounded":
    lits = bounded_degree_ksat_torch(n, cfg["k"], cfg["d"], 0xA111)
else:
    lits = uniform_ksat_torch(n, cfg["k"], int(cfg["m"] * a.scale), 0xA111)
m, k = lits.shape
torch.cuda.synchronize()
s = capi.Solver(device=0, sweep_smem_bytes=a.smem, flags=a.flags)
s.upload_fixedk_device(n, m, k, lits.data_ptr())
s.randomize(1)
ms, nv = s.time_sweep(a.reps)
alg = 4 * k * m + n // 8
out = dict(workload=a.workload, n=n, m=m, k=k, layout=s.layout_info(), sweep_ms=ms, n_violated=nv,
           achieved_GBps=alg / (ms * 1e-3) / 1e9)
for i in range(a.solves):
    s.randomize(10 + i)
    st = s.solve(10 + i, a.max_rounds)
    out[f"solve{i}"] = dict(ms=st.solve_ms, sweep_ms=st.sweep_ms, between_ms=st.between_sweeps_ms, iters=st.n_iterations, luby=st.n_luby_steps,
                            launches=st.n_kernel_launches, status=st.status)
print(json.dumps(out))
