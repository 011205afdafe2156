"""Clause-range sharded solve of one large instance over several GPUs (BASELINE config 4).

One process per GPU.  Every rank holds a contiguous clause range (global clause ids preserved) and a full replica
of the bit-packed assignment.  Per round:

    1. every rank sweeps its range and exports its violated clauses as records {global id, k literals};
    2. ``all_gather`` of the record counts, then of the records padded to the largest count
       (``torch.distributed``: NCCL over NVLink on GPUs, gloo in the CPU tests);
    3. every rank runs the identical independent-set + resample step on the full violated set.  Priorities and
       resample bits are Philox functions of (seed, round, global clause id / variable id), so the replicas stay
       bit-identical with no second exchange and the trajectory equals the single-GPU one for the same seed.

The round loop, partitioning and exchange live here and are backend-agnostic: ``CudaShardBackend`` drives the C ABI
(``alll_shard_sweep`` / ``alll_shard_round``); the CPU tests plug in a backend built on the oracle to exercise the
same host logic under gloo.  The reference has no counterpart: it is single-process, shared-memory OpenMP.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch
import torch.distributed as dist


def partition(m: int, world: int) -> list[tuple[int, int]]:
    """Contiguous, balanced clause ranges: rank r owns [lo_r, hi_r); sizes differ by at most one."""
    base, rem = divmod(m, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


@dataclass
class ShardedStats:
    n_iterations: int
    n_resamples: int
    avg_mis_size: int
    sum_mis_size: int
    n_clause_evals: int          # m_global * n_iterations
    status: int                  # 0 OK, 1 MAX_ROUNDS
    solve_ms: float              # device-clock interval of the round loop, max over ranks
    trace_u: list
    trace_s: list
    n_incremental_rounds: int = 0


class CudaShardBackend:
    """One ``alll_handle`` on this rank's GPU; record buffers are torch tensors (device memory is plumbing)."""

    def __init__(self, device: int):
        from . import capi

        self.capi = capi
        self.device = torch.device("cuda", device)
        self.solver = capi.Solver(device=device)
        self.k = 0
        self.send = None

    @property
    def comm_device(self):
        return self.device

    def upload(self, n_vars: int, lits_local, id_base: int):
        """``lits_local``: (m_local, k) uint32 numpy array or int32 CUDA tensor with this rank's clause range."""
        if isinstance(lits_local, torch.Tensor):
            m, k = lits_local.shape
            self.solver.upload_fixedk_device(n_vars, int(m), int(k), lits_local.data_ptr())
        else:
            m, k = lits_local.shape
            self.solver.upload_fixedk(n_vars, lits_local)
        self.solver.set_id_base(id_base)
        self.solver.reset_stats()
        self.k, self.m_local = int(k), int(m)
        cap = max(4096, int(self.m_local * 2.0 ** (-self.k) * 2) + 1024)
        self.send = torch.empty((cap, self.k + 1), dtype=torch.int32, device=self.device)

    def randomize(self, seed: int):
        self.solver.randomize(seed)

    def set_assignment(self, bools):
        self.solver.set_assignment(bools)

    def get_assignment(self):
        return self.solver.get_assignment()

    def sweep_export(self):
        """Returns (records tensor [cap, k+1] on the device, n_local)."""
        while True:
            try:
                n = self.solver.shard_sweep(self.send.data_ptr(), self.send.shape[0])
                return self.send, n
            except self.capi.AlllError as e:
                if e.status != self.capi.CAPACITY:
                    raise
                self.send = torch.empty((self.send.shape[0] * 4, self.k + 1), dtype=torch.int32, device=self.device)

    def shard_round(self, recs: torch.Tensor, counts, seed: int, rnd: int):
        """``recs``: [R, cap, k+1] gathered records on the device."""
        torch.cuda.current_stream(self.device).synchronize()      # the NCCL all-gather ran on torch's stream
        return self.solver.shard_round(recs.data_ptr(), counts, recs.shape[1], seed, rnd)

    def stats(self):
        return self.solver.get_stats()

    def clock(self):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record(torch.cuda.current_stream(self.device))
        return ev

    @staticmethod
    def elapsed_ms(a, b):
        b.synchronize()
        return a.elapsed_time(b)


class ShardedSolver:
    """Backend-agnostic driver: partition, per-round exchange, termination, statistics."""

    def __init__(self, backend, rank: int = 0, world: int = 1, group=None):
        self.backend, self.rank, self.world, self.group = backend, rank, world, group
        self.m_global = 0
        self.k = 0

    # -- upload -------------------------------------------------------------------------------------------
    def upload_range(self, n_vars: int, lits_local, m_global: int, id_base: int):
        """This rank already holds its own clause range."""
        self.n_vars, self.m_global, self.k = n_vars, int(m_global), int(lits_local.shape[1])
        self.backend.upload(n_vars, lits_local, id_base)

    def upload_full(self, n_vars: int, lits_full):
        """Every rank sees the whole (m, k) literal matrix and keeps only its range."""
        lo, hi = partition(int(lits_full.shape[0]), self.world)[self.rank]
        self.upload_range(n_vars, lits_full[lo:hi], int(lits_full.shape[0]), lo)

    # -- exchange -----------------------------------------------------------------------------------------
    def _all_gather_records(self, send: torch.Tensor, n_local: int):
        dev = send.device
        if self.world == 1:
            cap = max(n_local, 1)
            return send[:cap].unsqueeze(0), [n_local]
        mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
        counts_t = torch.empty(self.world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(counts_t, mine, group=self.group)
        counts = [int(c) for c in counts_t.tolist()]
        cap = max(max(counts), 1)
        if send.shape[0] < cap:                                   # another rank has more violated clauses than our buffer
            grown = torch.empty((cap, send.shape[1]), dtype=send.dtype, device=dev)
            grown[: send.shape[0]] = send
            send = grown
        recv = torch.empty((self.world, cap, send.shape[1]), dtype=send.dtype, device=dev)
        dist.all_gather_into_tensor(recv.view(-1), send[:cap].contiguous().view(-1), group=self.group)
        return recv, counts

    # -- round loop ---------------------------------------------------------------------------------------
    def solve(self, seed: int, max_rounds: int = 1 << 20) -> ShardedStats:
        be = self.backend
        t0 = be.clock()
        status, rnd = 1, 0
        n_iter = n_res = sum_mis = 0
        trace_u, trace_s = [], []
        max_rounds = max(max_rounds, 1)
        while rnd < max_rounds:
            send, n_local = be.sweep_export()
            recs, counts = self._all_gather_records(send, n_local)
            n_total, n_s, n_r = be.shard_round(recs, counts, seed, rnd)
            n_iter += 1                                          # every sweep counts, also the terminal one (SATInstance.h:261)
            trace_u.append(n_total)
            trace_s.append(n_s)
            if n_total == 0:                                     # SATInstance.h:285-287
                status = 0
                break
            sum_mis += n_s
            n_res += n_r
            rnd += 1
        t1 = be.clock()
        ms = be.elapsed_ms(t0, t1)
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=be.comm_device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            ms = float(t.item())
        return ShardedStats(n_iter, n_res, sum_mis // n_iter, sum_mis, self.m_global * n_iter, status, ms, trace_u, trace_s)


class P2PShardedSolver:
    """Sharded solve with the exchange fused into the kernels (NVLink P2P stores through CUDA IPC mappings).

    The round loop runs inside ``alll_solve_p2p`` with no NCCL call and no host round trip per round;
    ``torch.distributed`` is only used once to exchange the 64-byte IPC handles and as a barrier between solves.
    Needs one process per GPU on one node (peer access between all GPUs)."""

    def __init__(self, device: int, rank: int, world: int, group=None, persistent: bool = False, flags: int = 0):
        from . import capi

        self.capi = capi
        # persistent: every rank runs its whole solve as ONE cooperative kernel (ALLL_FLAG_P2P_PERSISTENT); only valid
        # when every rank has a GPU of its own -- the kernels of all ranks must be resident at the same time.
        # flags: e.g. capi.FLAG_INCREMENTAL (each rank then walks the occurrence lists of its own clause range)
        self.solver = capi.Solver(device=device, flags=(capi.FLAG_P2P_PERSISTENT if persistent else 0) | flags)
        self.device = torch.device("cuda", device)
        self.rank, self.world, self.group = rank, world, group
        self.epoch = 0
        self.m_global = 0

    def upload_range(self, n_vars: int, lits_local, m_global: int, id_base: int, cap_records: int | None = None):
        if isinstance(lits_local, torch.Tensor):
            m, k = lits_local.shape
            self.solver.upload_fixedk_device(n_vars, int(m), int(k), lits_local.data_ptr())
        else:
            m, k = lits_local.shape
            self.solver.upload_fixedk(n_vars, lits_local)
        self.solver.set_id_base(id_base)
        self.m_global = int(m_global)
        if cap_records is None:                      # a quarter of the widest clause range, with some headroom
            cap_records = (self.m_global // self.world) // 4 + 8192
        mine = self.solver.p2p_create(self.world, self.rank, int(cap_records))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, mine, group=self.group)
        else:
            handles = [mine]
        self.solver.p2p_connect(handles)
        if self.world > 1:
            dist.barrier(group=self.group)

    def randomize(self, seed: int):
        self.solver.randomize(seed)

    def get_assignment(self):
        return self.solver.get_assignment()

    def solve(self, seed: int, max_rounds: int = 1 << 19):
        if self.world > 1:
            dist.barrier(group=self.group)          # nobody may still be reading the previous solve's rounds
        self.epoch += 1
        st = self.solver.solve_p2p(seed, self.m_global, self.epoch, max_rounds)
        ms = st.solve_ms
        # incremental rounds evaluate only the clauses next to resampled variables: every rank counted its own range
        full = self.m_global * (st.n_iterations - st.n_incremental_rounds)
        evals = st.n_clause_evals
        if self.world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=self.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            ms = float(t.item())
            if st.n_incremental_rounds:
                e = torch.tensor([st.n_clause_evals - full], dtype=torch.int64, device=self.device)
                dist.all_reduce(e, op=dist.ReduceOp.SUM, group=self.group)
                evals = full + int(e.item())
        return ShardedStats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, evals,
                            st.status, ms, [], [], st.n_incremental_rounds)


# ---- portfolio and batched instances over the GPUs of one box (BASELINE config 5) -----------

def partition_round_robin(n_items: int, world: int, rank: int) -> np.ndarray:
    """Seed (or instance) s runs on GPU s mod world."""
    return np.arange(rank, n_items, world, dtype=np.int64)


class CudaPortfolioBackend:
    """Per-rank compute of the multi-GPU portfolio on the C ABI: one CTA per seed, one winner word for ALL ranks
    (device memory of rank 0, peer-mapped through CUDA IPC, claimed with a system-scope atomicCAS)."""

    def __init__(self, device: int):
        from . import capi

        self.solver = capi.Solver(device=device)
        self.comm_device = torch.device("cuda", device)

    def upload(self, n_vars: int, lits: np.ndarray):
        lits = np.ascontiguousarray(lits, np.uint32)
        self.solver.batch_upload(n_vars, lits.shape[1], np.array([0, lits.shape[0]], np.uint64), lits)

    def flag_create(self) -> bytes:
        return self.solver.flag_create()

    def flag_open(self, handle: bytes):
        self.solver.flag_open(handle)

    def flag_reset(self):
        self.solver.flag_reset()

    def run(self, seeds: np.ndarray, job_base: int, max_rounds: int):
        """-> (status per local job, assignment of the local winner or None, device ms)"""
        self.solver.batch_set_job_base(job_base)
        stats, assign, winner, ms = self.solver.batch_solve(seeds, max_rounds=max_rounds, portfolio=2, want_assignments=True)
        won = np.flatnonzero(stats["status"] == 0)
        return stats["status"].copy(), (assign[won[0]] if len(won) else None), ms


class MultiGpuPortfolio:
    """One instance, ``n_seeds`` solver seeds spread over all ranks (seed s -> rank s mod world); the first job
    anywhere to satisfy every clause claims the shared winner word and every other job stops at its next round.

    ``torch.distributed`` only moves the 64-byte flag handle once, acts as the barrier around a portfolio and
    collects who won; during the solve the ranks interact through the flag word alone."""

    def __init__(self, backend, rank: int = 0, world: int = 1, group=None):
        self.be, self.rank, self.world, self.group = backend, rank, world, group
        handle = [self.be.flag_create() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(handle, src=0, group=group)
            if rank != 0:
                self.be.flag_open(handle[0])
            dist.barrier(group=group)

    def upload(self, n_vars: int, lits):
        self.n_vars = n_vars
        self.be.upload(n_vars, lits)

    def solve(self, seeds, max_rounds: int = 1 << 20):
        """-> dict(winner_seed_index, winner_rank, assignment (on every rank), ms (max over ranks), n_finished)"""
        seeds = np.ascontiguousarray(seeds, np.uint64)
        mine = partition_round_robin(len(seeds), self.world, self.rank)
        if self.world > 1:
            dist.barrier(group=self.group)               # nobody is still polling the previous portfolio's word
        if self.rank == 0:
            self.be.flag_reset()
        if self.world > 1:
            dist.barrier(group=self.group)
        status, assign, ms = self.be.run(seeds[mine], job_base=self.rank * len(seeds), max_rounds=max_rounds)
        won = np.flatnonzero(status == 0)
        # exactly one job in the world may hold status OK: the one whose compare-and-swap found the word open
        info = torch.tensor([int(mine[won[0]]) if len(won) else -1, len(won), ms], dtype=torch.float64, device=self.be.comm_device)
        if self.world > 1:
            allinfo = [torch.zeros_like(info) for _ in range(self.world)]
            dist.all_gather(allinfo, info, group=self.group)
        else:
            allinfo = [info]
        allinfo = [t.cpu().numpy() for t in allinfo]
        n_finished = int(sum(t[1] for t in allinfo))
        winner_rank = next((r for r, t in enumerate(allinfo) if t[1] > 0), -1)
        winner_seed = int(allinfo[winner_rank][0]) if winner_rank >= 0 else -1
        buf = torch.zeros(self.n_vars, dtype=torch.uint8, device=self.be.comm_device)
        if winner_rank == self.rank and assign is not None:
            buf = torch.from_numpy(np.ascontiguousarray(assign)).to(self.be.comm_device)
        if self.world > 1 and winner_rank >= 0:
            dist.broadcast(buf, src=winner_rank, group=self.group)
        return dict(winner_seed_index=winner_seed, winner_rank=winner_rank, n_finished=n_finished,
                    assignment=buf.cpu().numpy() if winner_rank >= 0 else None, ms=float(max(t[2] for t in allinfo)))
