#!/usr/bin/env python
"""One process, N GPUs (alll_multi_*): clause-range sharded solve of a BASELINE workload, device-timed, optionally with
the in-kernel phase stamps of every rank (ALLL_TRACE=1) and the host-side stage times (ALLL_TRACE_HOST=1).

    python tools/prof_multi.py --gpus 8 [--workload cfg4] [--solves 5] [--incremental] [--e2e 3]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from alllsatisfiabilitysolver_b200 import capi  # noqa: E402
from alllsatisfiabilitysolver_b200.instances import CONFIGS, bounded_degree_ksat_torch, uniform_ksat_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--solves", type=int, default=5)
ap.add_argument("--incremental", action="store_true")
ap.add_argument("--e2e", type=int, default=0)
ap.add_argument("--both", action="store_true", help="full sweeps, then incremental, on the same instance")
a = ap.parse_args()
cfg = CONFIGS[a.workload]
n = int(cfg["n"] * a.scale)
lits = (bounded_degree_ksat_torch(n, cfg["k"], cfg["d"], 0xA115) if cfg["kind"] == "bounded"
        else uniform_ksat_torch(n, cfg["k"], int(cfg["m"] * a.scale), 0xA115))
m, k = int(lits.shape[0]), int(lits.shape[1])
host_t = torch.empty(lits.shape, dtype=lits.dtype, pin_memory=True)
host_t.copy_(lits)
torch.cuda.synchronize()
del lits
torch.cuda.empty_cache()
host = host_t.numpy().view(np.uint32)
modes = [False, True] if a.both else [a.incremental]
for inc in modes:
  out = dict(gpus=a.gpus, workload=a.workload, n=n, m=m, k=k, incremental=inc, solves=[])
  with capi.MultiSolver(list(range(a.gpus)), flags=capi.FLAG_INCREMENTAL if inc else 0) as ms:
      t0 = time.perf_counter()
      ms.upload_fixedk(n, host)
      out["first_upload_ms"] = (time.perf_counter() - t0) * 1e3
      out["layout"] = ms.info()
      for i in range(a.solves):
          ms.randomize(100 + i)
          st = ms.solve(100 + i)
          out["solves"].append(dict(ms=st.solve_ms, sweep_ms=st.sweep_ms, between_ms=st.between_sweeps_ms, iters=st.n_iterations,
                                    incr_rounds=st.n_incremental_rounds, status=st.status))
      out["verified"] = ms.verify()
      if a.e2e:
          vout = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
          ts = []
          for i in range(-1, a.e2e):
              t0 = time.perf_counter()
              ms.upload_fixedk(n, host)
              t1 = time.perf_counter()
              ms.randomize(200 + i)
              st = ms.solve(200 + i)
              t2 = time.perf_counter()
              ms.get_assignment(vout)
              t3 = time.perf_counter()
              if i >= 0:
                  ts.append(((t3 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
          out["e2e_ms"] = dict(total=float(np.mean([x[0] for x in ts])), upload=float(np.mean([x[1] for x in ts])),
                               randomize_solve=float(np.mean([x[2] for x in ts])), readback=float(np.mean([x[3] for x in ts])))
  print(json.dumps(out), flush=True)
