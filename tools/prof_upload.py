#!/usr/bin/env python
"""Times the host-facing upload path (alll_upload_fixedk from pinned / pageable memory) and its device-only part."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from alllsatisfiabilitysolver_b200 import capi
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat_torch

n, k, d = 10_000_000, 8, 32
lits = bounded_degree_ksat_torch(n, k, d, 0xA111)
m = lits.shape[0]
pinned = torch.empty(lits.shape, dtype=lits.dtype, pin_memory=True); pinned.copy_(lits); torch.cuda.synchronize()
pageable = pinned.numpy().copy().view(np.uint32)
pin_np = pinned.numpy().view(np.uint32)
s = capi.Solver(device=0)
out = {}
for name, fn in [("device", lambda: s.upload_fixedk_device(n, m, k, lits.data_ptr())),
                 ("pinned", lambda: s.upload_fixedk(n, pin_np)),
                 ("pageable", lambda: s.upload_fixedk(n, pageable))]:
    ts = []
    for _ in range(3):
        t = time.perf_counter(); fn(); ts.append((time.perf_counter() - t) * 1e3)
    out[name] = ts
t = time.perf_counter(); s.randomize(1); st = s.solve(1); a = s.get_assignment(); out["solve+get_ms"] = (time.perf_counter() - t) * 1e3
t = time.perf_counter(); a = s.get_assignment(); out["get_ms"] = (time.perf_counter() - t) * 1e3
t = time.perf_counter(); s.set_assignment(a); out["set_ms"] = (time.perf_counter() - t) * 1e3
print(json.dumps(out))
