#!/usr/bin/env python
"""A/B of the packed H2D transport (ALLL_H2D_PACK) on one B200: the bench's end-to-end step -- alll_upload_fixedk(host) +
randomize + solve + get_assignment -- from page-locked and from pageable caller memory, packed against plain.

    python tools/prof_upload_pack.py [--workload cfg4] [--steps 5]
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
from alllsatisfiabilitysolver_b200.instances import CONFIGS, bounded_degree_ksat_torch  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--pageable-steps", type=int, default=2)
ap.add_argument("--chunk-rows", default="0", help="comma list of ALLL_H2D_CHUNK_ROWS values (0 = the library's 64 MB chunks)")
a = ap.parse_args()
cfg = CONFIGS[a.workload]
n, k = cfg["n"], cfg["k"]
lits_t = bounded_degree_ksat_torch(n, k, cfg["d"], 0xA111 + 4)
m = int(lits_t.shape[0])
pinned_t = torch.empty(lits_t.shape, dtype=lits_t.dtype, pin_memory=True)
pinned_t.copy_(lits_t)
torch.cuda.synchronize()
pinned = pinned_t.numpy().view(np.uint32)
pageable = np.array(pinned, copy=True)
out_t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
out = out_t.numpy()
res = {"workload": a.workload, "n": n, "m": m, "k": k, "host_buffer_bytes": 4 * k * m, "host_threads": os.cpu_count()}
ref_assign = None
for src_name, src, steps in (("page_locked", pinned, a.steps), ("pageable", pageable, a.pageable_steps)):
  if steps <= 0:
    continue
  for chunk_rows in [int(x) for x in a.chunk_rows.split(",")]:
    if chunk_rows:
        os.environ["ALLL_H2D_CHUNK_ROWS"] = str(chunk_rows)
    else:
        os.environ.pop("ALLL_H2D_CHUNK_ROWS", None)
    for mode in ("0", "1", "0", "1"):                       # twice each, interleaved
        os.environ["ALLL_H2D_PACK"] = mode
        s = capi.Solver(device=0)

        def step(i):
            s.upload_fixedk(n, src)
            s.randomize(2000 + i)
            st = s.solve(2000 + i, 2000)
            s.get_assignment(out)
            return st

        step(-1)
        torch.cuda.synchronize()
        ts = []
        up_ms = []
        for i in range(steps):
            t0 = time.perf_counter()
            st = step(i)
            ts.append((time.perf_counter() - t0) * 1e3)
            t0 = time.perf_counter()
            s.upload_fixedk(n, src)
            up_ms.append((time.perf_counter() - t0) * 1e3)
        assert st.status == 0
        s.randomize(7)
        s.solve(7, 2000)
        got = s.get_assignment().copy()
        if ref_assign is None:
            ref_assign = got
        same = bool(np.array_equal(got, ref_assign))
        key = f"{src_name}_chunk{chunk_rows}_pack{mode}"
        rec = {"e2e_ms": ts, "upload_only_ms": up_ms, "upload_info": s.upload_info(), "same_assignment_as_first_variant": same, "verified": bool(s.verify())}
        res.setdefault(key, []).append(rec)
        print(key, "e2e", ["%.2f" % x for x in ts], "upload", ["%.2f" % x for x in up_ms], rec["upload_info"], same, file=sys.stderr)
        s.close()
print(json.dumps(res))
