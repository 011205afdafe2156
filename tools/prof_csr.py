#!/usr/bin/env python
"""Driver for ncu / timing of the warp-cooperative CSR sweep: a ragged instance (widths wmin..wmax), kept in CSR form.

    python tools/prof_csr.py [--m 40000000] [--n 10000000] [--wmin 3] [--wmax 8] [--reps 5] [--solves 0] [--smem BYTES]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from alllsatisfiabilitysolver_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--m", type=int, default=40_000_000)
ap.add_argument("--n", type=int, default=10_000_000)
ap.add_argument("--wmin", type=int, default=3)
ap.add_argument("--wmax", type=int, default=8)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--solves", type=int, default=0)
ap.add_argument("--smem", type=int, default=0)
ap.add_argument("--flags", type=int, default=0)
a = ap.parse_args()
g = torch.Generator(device="cuda")
g.manual_seed(0xA11A)
widths = torch.randint(a.wmin, a.wmax + 1, (a.m,), generator=g, device="cuda", dtype=torch.int64)
off = torch.zeros(a.m + 1, dtype=torch.int64, device="cuda")
off[1:] = torch.cumsum(widths, 0)
L = int(off[-1])
lit = (torch.randint(0, a.n, (L,), generator=g, device="cuda", dtype=torch.int32) * 2 +
       torch.randint(0, 2, (L,), generator=g, device="cuda", dtype=torch.int32))
off_np, lit_np = off.cpu().numpy().astype(np.uint64), lit.cpu().numpy().view(np.uint32)
del widths, off, lit
torch.cuda.empty_cache()
s = capi.Solver(device=0, sweep_smem_bytes=a.smem, flags=capi.FLAG_FORCE_CSR | a.flags)
s.upload_csr(a.n, off_np, lit_np)
s.randomize(1)
ms, nv = s.time_sweep(a.reps)
alg = 4 * L + 8 * (a.m + 1)
out = dict(m=a.m, n=a.n, L=L, layout=s.layout_info(), sweep_ms=ms, n_violated=nv, algorithmic_GBps=alg / (ms * 1e-3) / 1e9,
           read_GBps=s.layout_info()["literal_bytes"] / (ms * 1e-3) / 1e9)
for i in range(a.solves):
    s.randomize(10 + i)
    st = s.solve(10 + i, 300)
    out[f"solve{i}"] = dict(ms=st.solve_ms, sweep_ms=st.sweep_ms, between_ms=st.between_sweeps_ms, iters=st.n_iterations, status=st.status,
                            launches=st.n_kernel_launches)
print(json.dumps(out))
