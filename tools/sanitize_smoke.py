#!/usr/bin/env python
"""Small run through every kernel for compute-sanitizer (memcheck / racecheck / synccheck, one tool per call):

    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from alllsatisfiabilitysolver_b200 import capi
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat, uniform_ksat

rng = np.random.default_rng(0)
def solve_case(n, lits, **cfg):
    with capi.Solver(**cfg) as s:
        s.upload_fixedk(n, lits)
        s.randomize(3)
        cnt, ids = s.eval()
        u, sset, r = s.round(3, 0)
        st = s.solve(3, 200)
        assert st.status in (0, 1)
        return st.n_iterations

l8 = bounded_degree_ksat(6000, 8, 32, 1)
print("resident k8", solve_case(6000, l8))
print("bucketed k8", solve_case(6000, l8, sweep_smem_bytes=256))
print("gather k8", solve_case(6000, l8, sweep_smem_bytes=256, flags=1))
print("k3", solve_case(3000, uniform_ksat(3000, 3, 6000, 2)))
print("k12 generic", solve_case(4000, (rng.integers(0, 4000, (5000, 12)) * 2 + rng.integers(0, 2, (5000, 12))).astype(np.uint32), sweep_smem_bytes=256))
# CSR
clauses = [list((rng.choice(500, size=int(rng.integers(2, 7)), replace=False) * 2 + rng.integers(0, 2)).astype(np.uint32)) for _ in range(400)]
off = np.zeros(401, np.uint64); off[1:] = np.cumsum([len(c) for c in clauses]); lit = np.array([x for c in clauses for x in c], np.uint32)
with capi.Solver() as s:
    s.upload_csr(500, off, lit); s.randomize(1); print("csr", s.solve(1, 200).n_iterations)
# large-U grid MIS path: all-positive disjoint clauses under the all-false assignment (|U| = m > 8192)
m = 20000
with capi.Solver() as s:
    s.upload_fixedk(3 * m, (np.arange(3 * m, dtype=np.uint32).reshape(m, 3)) * 2)
    s.set_assignment(np.zeros(3 * m, np.uint8)); u, sset, r = s.round(1, 0); print("grid mis", len(u), len(sset))
# batch + portfolio
insts = [bounded_degree_ksat(1500, 5, 3, 10 + i) for i in range(6)]
boff = np.zeros(7, np.uint64); boff[1:] = np.cumsum([x.shape[0] for x in insts])
with capi.Solver() as s:
    s.batch_upload(1500, 5, boff, np.concatenate(insts))
    st, a, w, ms = s.batch_solve(np.arange(6, dtype=np.uint64)); print("batch", st["n_iterations"].tolist())
    s.batch_upload(1500, 5, boff[:2], insts[0])
    st, a, w, ms = s.batch_solve(np.arange(32, dtype=np.uint64), portfolio=True); print("portfolio winner", w)
# sharded (two handles on one device)
import torch
from alllsatisfiabilitysolver_b200.sharded import CudaShardBackend, partition
bes = []
for lo, hi in partition(l8.shape[0], 2):
    b = CudaShardBackend(0); b.upload(6000, l8[lo:hi], lo); b.randomize(5); bes.append(b)
for rnd in range(50):
    sends = [b.sweep_export() for b in bes]; counts = [c for _, c in sends]; cap = max(max(counts), 1)
    recs = torch.zeros((2, cap, 9), dtype=torch.int32, device="cuda")
    for r, (t, c) in enumerate(sends): recs[r, :c] = t[:c]
    torch.cuda.synchronize()
    outs = [b.shard_round(recs, counts, 5, rnd) for b in bes]
    if outs[0][0] == 0: break
print("sharded rounds", rnd + 1)
print("SANITIZE_SMOKE_DONE")
