#!/usr/bin/env python
"""torchrun driver of the clause-range sharded solve (one process per GPU, NCCL).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_sharded.py [--workload cfg4] [--scale 1.0] [--solves 3] [--check]

Every rank generates only its own clause range?  No: the instance must be the same global object on every rank, so
rank 0's generator seed is shared and each rank generates the full instance on its GPU, then keeps its range
(generation is not timed).  --check compares the result with the single-GPU solve of the same seed on rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from alllsatisfiabilitysolver_b200 import capi  # noqa: E402
from alllsatisfiabilitysolver_b200.instances import CONFIGS, bounded_degree_ksat_torch, uniform_ksat_torch  # noqa: E402
from alllsatisfiabilitysolver_b200.sharded import CudaShardBackend, P2PShardedSolver, ShardedSolver, partition  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg4")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--solves", type=int, default=3)
ap.add_argument("--check", action="store_true")
ap.add_argument("--p2p", action="store_true", help="exchange fused into the kernels (NVLink P2P) instead of NCCL all-gather")
ap.add_argument("--persistent", action="store_true", help="with --p2p: one persistent solve kernel per rank")
ap.add_argument("--incremental", action="store_true", help="with --p2p --persistent: ALLL_FLAG_INCREMENTAL on every rank")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = CONFIGS[a.workload]
n = max(int(cfg["n"] * a.scale), 1000)
lits = (bounded_degree_ksat_torch(n, cfg["k"], cfg["d"], 0xA115) if cfg["kind"] == "bounded"
        else uniform_ksat_torch(n, cfg["k"], int(cfg["m"] * a.scale), 0xA115))
m, k = int(lits.shape[0]), int(lits.shape[1])
lo, hi = partition(m, world)[rank]
if a.p2p:
    ss = P2PShardedSolver(local, rank, world, persistent=a.persistent, flags=capi.FLAG_INCREMENTAL if a.incremental else 0)
    be = ss
else:
    be = CudaShardBackend(local)
    ss = ShardedSolver(be, rank, world)
ss.upload_range(n, lits[lo:hi].contiguous(), m, lo)
out = dict(world=world, n=n, m=m, k=k, mode=("p2p-persistent" if a.persistent else "p2p") if a.p2p else "nccl", solves=[])
for i in range(a.solves):
    if not a.p2p:
        be.solver.reset_stats()
    be.randomize(100 + i)
    st = ss.solve(100 + i)
    out["solves"].append(dict(ms=st.solve_ms, iters=st.n_iterations, resamples=st.n_resamples, status=st.status,
                              incremental_rounds=getattr(st, "n_incremental_rounds", 0),
                              clause_evals_per_s=st.n_clause_evals / (st.solve_ms * 1e-3)))
ok = True
if a.check:
    mine = torch.from_numpy(be.get_assignment()).cuda()
    if world > 1:
        ref = mine.clone()
        dist.broadcast(ref, 0)
        ok = bool((ref == mine).all())
    if rank == 0:
        s1 = capi.Solver(device=local)
        s1.upload_fixedk_device(n, m, k, lits.data_ptr())
        s1.randomize(100 + a.solves - 1)
        st1 = s1.solve(100 + a.solves - 1)
        ok = ok and np.array_equal(s1.get_assignment(), mine.cpu().numpy()) and s1.verify() and \
            (st1.n_iterations, st1.n_resamples) == (out["solves"][-1]["iters"], out["solves"][-1]["resamples"])
        out["single_gpu_ms"] = st1.solve_ms
    if world > 1:
        t = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok = bool(t.item())
out["ok"] = ok
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
