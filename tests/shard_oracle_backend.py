"""TEST-ONLY backend for ``ShardedSolver``: the per-rank compute done by the CPU oracle, tensors on the CPU.
Lets the world_size>1 host logic (partition, all-gather protocol, termination, statistics) run under gloo."""
import time

import numpy as np
import torch

from oracle.oracle import Oracle


class OracleShardBackend:
    comm_device = torch.device("cpu")

    def __init__(self, m_global: int):
        self.o = Oracle()
        self.m_global = m_global

    def upload(self, n_vars, lits_local, id_base):
        lits_local = np.ascontiguousarray(lits_local, np.uint32)
        self.n_vars, self.id_base = n_vars, id_base
        self.m_local, self.k = lits_local.shape
        self.off = np.arange(self.m_local + 1, dtype=np.uint64) * np.uint64(self.k)
        self.lit = lits_local.reshape(-1)
        self.lits2d = lits_local

    def randomize(self, seed):
        self.vars = self.o.randomize(self.n_vars, seed)

    def get_assignment(self):
        return self.vars.copy()

    def sweep_export(self):
        u = self.o.sweep(self.off, self.lit, self.vars)
        rec = np.empty((max(len(u), 1), self.k + 1), np.uint32)
        rec[: len(u), 0] = u + np.uint32(self.id_base)
        rec[: len(u), 1:] = self.lits2d[u]
        return torch.from_numpy(rec.view(np.int32)), len(u)

    def shard_round(self, recs, counts, seed, rnd):
        recs = recs.numpy().view(np.uint32)
        rows = [recs[r, : counts[r]] for r in range(len(counts))]
        allr = np.concatenate(rows, axis=0) if rows else np.empty((0, self.k + 1), np.uint32)
        n_total = allr.shape[0]
        if n_total == 0:
            return 0, 0, 0
        order = np.argsort(allr[:, 0], kind="stable")
        allr = allr[order]
        ids = np.ascontiguousarray(allr[:, 0])
        # CSR over the GLOBAL id space in which only the violated clauses have literals
        width = np.zeros(self.m_global, np.uint64)
        width[ids] = self.k
        off_g = np.zeros(self.m_global + 1, np.uint64)
        off_g[1:] = np.cumsum(width)
        lit_g = np.ascontiguousarray(allr[:, 1:].reshape(-1))
        s = self.o.priority_mis(self.n_vars, off_g, lit_g, ids, seed, rnd)
        n_r = self.o.resample(off_g, lit_g, s, seed, rnd, self.vars)
        return n_total, len(s), n_r

    def clock(self):
        return time.perf_counter()

    @staticmethod
    def elapsed_ms(a, b):
        return (b - a) * 1e3
