"""TEST-ONLY backend for ``MultiGpuPortfolio``: per-rank seeds solved by the CPU oracle; the "device flag" shared by the
ranks is a file claimed with O_CREAT|O_EXCL (atomic first-writer-wins, like the system-scope compare-and-swap)."""
import os
import time

import numpy as np
import torch

from oracle.oracle import Oracle, to_csr

OK, MAX_ROUNDS, PREEMPTED = 0, 1, 8


class OraclePortfolioBackend:
    comm_device = torch.device("cpu")

    def __init__(self, flag_dir: str):
        self.o = Oracle()
        self.flag_dir = flag_dir
        self.path = None

    def upload(self, n_vars, lits):
        self.n_vars = n_vars
        self.off, self.lit = to_csr(np.ascontiguousarray(lits, np.uint32))

    def flag_create(self) -> bytes:
        self.path = os.path.join(self.flag_dir, "winner.flag")
        return self.path.encode()

    def flag_open(self, handle: bytes):
        self.path = handle.decode()

    def flag_reset(self):
        if os.path.exists(self.path):
            os.remove(self.path)

    def run(self, seeds, job_base, max_rounds):
        t0 = time.perf_counter()
        status = np.full(len(seeds), PREEMPTED, np.int32)
        winner = None
        for j, seed in enumerate(seeds):
            if os.path.exists(self.path):                       # polled at the next "round boundary"
                break
            vars_ = self.o.randomize(self.n_vars, int(seed))
            st = self.o.solve(self.n_vars, self.off, self.lit, vars_, int(seed), max_rounds=max_rounds)
            if st.status != OK:
                status[j] = MAX_ROUNDS
                continue
            try:
                fd = os.open(self.path, os.O_CREAT | os.O_EXCL | os.O_WRONLY)
                os.write(fd, str(job_base + j).encode())
                os.close(fd)
                status[j] = OK
                winner = vars_
                break
            except FileExistsError:
                break
        return status, winner, (time.perf_counter() - t0) * 1e3
