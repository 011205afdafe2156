#!/usr/bin/env python
"""torchrun driver of the multi-GPU seed portfolio and of batched instances spread over the GPUs (BASELINE config 5).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/run_portfolio.py [--seeds 8192] [--instances 8192] [--n 10000] [--check]

Portfolio: one 5-SAT instance, seed s on GPU s mod N, ONE winner word for all GPUs (device memory of rank 0 mapped by
the others through CUDA IPC, claimed with a system-scope atomicCAS).  Batch: instance i on GPU i mod N, no exchange.
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
from alllsatisfiabilitysolver_b200.instances import CONFIGS, INSTANCE_SEED_BASE, bounded_degree_batch_torch, bounded_degree_ksat  # noqa: E402
from alllsatisfiabilitysolver_b200.sharded import CudaPortfolioBackend, MultiGpuPortfolio, partition_round_robin  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--seeds", type=int, default=8192)
ap.add_argument("--instances", type=int, default=8192)
ap.add_argument("--n", type=int, default=CONFIGS["cfg5"]["n"])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
c5 = CONFIGS["cfg5"]
out = {"world": world}

# ---- portfolio: same instance on every rank
lits = bounded_degree_ksat(a.n, c5["k"], c5["d"], seed=INSTANCE_SEED_BASE + 5)
pf = MultiGpuPortfolio(CudaPortfolioBackend(local), rank, world)
pf.upload(a.n, lits)
runs = []
for r in range(a.reps):
    res = pf.solve(np.arange(r * a.seeds, (r + 1) * a.seeds, dtype=np.uint64))
    ok = res["n_finished"] == 1 and res["assignment"] is not None
    if ok:
        v = res["assignment"].astype(bool)
        val = v[lits >> 1] ^ (lits & 1).astype(bool)           # literal true iff value != neg
        ok = bool(val.any(axis=1).all())
    runs.append({"ms": res["ms"], "winner_seed_index": res["winner_seed_index"], "winner_rank": res["winner_rank"],
                 "n_finished": res["n_finished"], "verified": ok})
out["portfolio"] = {"seeds": a.seeds, "runs": runs, "first_sat_ms": min(x["ms"] for x in runs), "all_verified": all(x["verified"] for x in runs)}

# ---- batch: instances round-robin over the ranks, no exchange
mine = partition_round_robin(a.instances, world, rank)
off, blits = bounded_degree_batch_torch(len(mine), a.n, c5["k"], c5["d"], INSTANCE_SEED_BASE + 5 + 1000 * rank)
s = capi.Solver(device=local)
s.batch_upload(a.n, c5["k"], off.numpy().astype(np.uint64), blits.cpu().numpy().view(np.uint32))
best, solved = None, 0
for r in range(a.reps):
    if world > 1:
        dist.barrier()
    st, _, _, ms = s.batch_solve(mine.astype(np.uint64) + np.uint64(r * a.instances), want_assignments=False)
    t = torch.tensor([ms, float((st["status"] == 0).sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        ms, solved = float(tmax[0]), int(tsum[1])
    else:
        solved = int(t[1])
    best = ms if best is None else min(best, ms)
out["batch"] = {"instances": a.instances, "batch_ms": best, "instances_per_sec": a.instances / (best * 1e-3), "solved": solved}
out["ok"] = bool(out["portfolio"]["all_verified"] and solved == a.instances)
if rank == 0:
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if (out["ok"] or not a.check) else 1)
