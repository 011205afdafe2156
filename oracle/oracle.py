"""TEST INFRASTRUCTURE ONLY -- ctypes loaders for the two CPU checkers.

* ``Oracle``   : oracle/liballl_oracle.so, the plain-C restatement (alll_oracle.c).
* ``Reference``: oracle/_ref/liballl_ref.so, the UNMODIFIED reference headers behind
  a thin extern "C" shim (ref_harness.cpp), built by oracle/Makefile from the sources
  where they lie under /root/reference.  It is prebuilt here and travels to the GPU box.

Nothing here reads /root/reference at run time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liballl_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "liballl_ref.so")

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")


def build(ref: bool = True) -> None:
    """Compile the checkers (building the checker is not using it)."""
    targets = ["oracle"] + (["ref"] if ref else [])
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def have_reference() -> bool:
    return os.path.exists(_REF_SO)


def to_csr(lits) -> tuple[np.ndarray, np.ndarray]:
    """(m,k) uint32 matrix, or (off, lit) pair, or list of lists -> (off u64, lit u32)."""
    if isinstance(lits, tuple):
        off, lit = lits
        return np.ascontiguousarray(off, np.uint64), np.ascontiguousarray(lit, np.uint32)
    if isinstance(lits, np.ndarray) and lits.ndim == 2:
        m, k = lits.shape
        off = (np.arange(m + 1, dtype=np.uint64) * np.uint64(k)).astype(np.uint64)
        return off, np.ascontiguousarray(lits.reshape(-1), np.uint32)
    off = np.zeros(len(lits) + 1, np.uint64)
    off[1:] = np.cumsum([len(c) for c in lits], dtype=np.uint64)
    flat = np.array([l for c in lits for l in c], dtype=np.uint32)
    return off, flat


@dataclass
class Stats:
    n_iterations: int
    n_resamples: int
    avg_mis_size: int
    sum_mis_size: int = 0
    n_clause_evals: int = 0
    status: int = 0
    seconds: float = 0.0


class _StatsC(C.Structure):
    _fields_ = [("n_iterations", C.c_uint64), ("n_resamples", C.c_uint64),
                ("avg_mis_size", C.c_uint64), ("sum_mis_size", C.c_uint64),
                ("n_clause_evals", C.c_uint64)]


class Oracle:
    """Plain-C restatement (oracle/alll_oracle.c)."""

    def __init__(self):
        if not os.path.exists(_ORACLE_SO):
            build(ref=False)
        L = self.lib = C.CDLL(_ORACLE_SO)
        L.alll_oracle_sweep.restype = C.c_uint64
        L.alll_oracle_sweep.argtypes = [C.c_uint64, _u64p, _u32p, _u8p, C.c_void_p]
        L.alll_oracle_verify.restype = C.c_int
        L.alll_oracle_verify.argtypes = [C.c_uint64, _u64p, _u32p, _u8p]
        L.alll_oracle_dependent.restype = C.c_int
        L.alll_oracle_dependent.argtypes = [_u32p, C.c_uint64, _u32p, C.c_uint64]
        L.alll_oracle_batches.argtypes = [C.c_uint64, C.c_int, _u16p]
        L.alll_oracle_greedy_mis.restype = C.c_uint64
        L.alll_oracle_greedy_mis.argtypes = [C.c_uint64, _u64p, _u32p, _u32p, C.c_uint64, C.c_int, _u32p]
        L.alll_oracle_check_signed.restype = C.c_int
        L.alll_oracle_check_signed.argtypes = [C.c_int64, _i32p, _i32p, _u8p]
        L.alll_oracle_philox4x32_10.argtypes = [_u32p, _u32p, _u32p]
        L.alll_oracle_random_bit.restype = C.c_uint32
        L.alll_oracle_random_bit.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32]
        L.alll_oracle_priority.restype = C.c_uint32
        L.alll_oracle_priority.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.alll_oracle_randomize.argtypes = [C.c_uint64, C.c_uint64, _u8p]
        L.alll_oracle_gen_materialize.restype = C.c_int
        L.alll_oracle_gen_materialize.argtypes = [C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint32, _u32p]
        L.alll_oracle_priority_mis.restype = C.c_uint64
        L.alll_oracle_priority_mis.argtypes = [C.c_uint64, _u64p, _u32p, _u32p, C.c_uint64,
                                               C.c_uint64, C.c_uint32, _u8p, _u32p]
        L.alll_oracle_resample.restype = C.c_uint64
        L.alll_oracle_resample.argtypes = [_u64p, _u32p, _u32p, C.c_uint64, C.c_uint64, C.c_uint32, _u8p]
        L.alll_oracle_round.restype = C.c_uint64
        L.alll_oracle_round.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, _u8p, C.c_uint64, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        L.alll_oracle_solve.restype = C.c_int
        L.alll_oracle_solve.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, _u8p, C.c_uint64, C.c_uint64,
                                        C.POINTER(_StatsC), C.c_void_p, C.c_void_p]
        L.alll_oracle_solve_greedy.restype = C.c_int
        L.alll_oracle_solve_greedy.argtypes = [C.c_uint64, C.c_uint64, _u64p, _u32p, _u8p, C.c_uint64,
                                               C.c_int, C.c_uint64, C.POINTER(_StatsC)]

    # -- reference restatements ------------------------------------------------
    def sweep(self, off, lit, vars_) -> np.ndarray:
        m = len(off) - 1
        out = np.empty(max(m, 1), np.uint32)
        n = self.lib.alll_oracle_sweep(m, off, lit, vars_, out.ctypes.data)
        return out[:n].copy()

    def verify(self, off, lit, vars_) -> bool:
        return bool(self.lib.alll_oracle_verify(len(off) - 1, off, lit, vars_))

    def dependent(self, l1, l2) -> bool:
        l1 = np.ascontiguousarray(l1, np.uint32)
        l2 = np.ascontiguousarray(l2, np.uint32)
        return bool(self.lib.alll_oracle_dependent(l1, len(l1), l2, len(l2)))

    def batches(self, m: int, n_threads: int) -> np.ndarray:
        out = np.empty(max(m, 1), np.uint16)
        self.lib.alll_oracle_batches(m, n_threads, out)
        return out[:m]

    def greedy_mis(self, off, lit, u_ids, n_threads: int) -> np.ndarray:
        u_ids = np.ascontiguousarray(u_ids, np.uint32)
        out = np.empty(max(len(u_ids), 1), np.uint32)
        n = self.lib.alll_oracle_greedy_mis(len(off) - 1, off, lit, u_ids, len(u_ids), n_threads, out)
        return out[:n].copy()

    def check_signed(self, l_c_num, l_val, v_val) -> bool:
        l_c_num = np.ascontiguousarray(l_c_num, np.int32)
        l_val = np.ascontiguousarray(l_val, np.int32)
        return bool(self.lib.alll_oracle_check_signed(len(l_c_num), l_c_num, l_val,
                                                      np.ascontiguousarray(v_val, np.uint8)))

    # -- deterministic round specification ------------------------------------------
    def philox(self, ctr, key) -> np.ndarray:
        out = np.zeros(4, np.uint32)
        self.lib.alll_oracle_philox4x32_10(np.ascontiguousarray(ctr, np.uint32),
                                           np.ascontiguousarray(key, np.uint32), out)
        return out

    def random_bit(self, seed, stream, rnd, v) -> int:
        return int(self.lib.alll_oracle_random_bit(seed, stream, rnd, v))

    def priority(self, seed, rnd, c) -> int:
        return int(self.lib.alll_oracle_priority(seed, rnd, c))

    def randomize(self, n_vars: int, seed: int) -> np.ndarray:
        out = np.zeros(max(n_vars, 1), np.uint8)
        self.lib.alll_oracle_randomize(n_vars, seed, out)
        return out[:n_vars]

    def gen_materialize(self, kind: int, n_vars: int, m: int, k: int, seed: int, d: int = 0) -> np.ndarray:
        """All clauses of a built-in enumerated instance as an (m, k) literal matrix."""
        out = np.zeros((max(m, 1), k), np.uint32)
        if self.lib.alll_oracle_gen_materialize(kind, n_vars, m, k, seed, d, out.reshape(-1)) != 0:
            raise ValueError("generator parameters outside the specification")
        return out[:m]

    def priority_mis(self, n_vars, off, lit, u_ids, seed, rnd) -> np.ndarray:
        u_ids = np.ascontiguousarray(u_ids, np.uint32)
        scratch = np.zeros(max(n_vars, 1), np.uint8)
        out = np.empty(max(len(u_ids), 1), np.uint32)
        n = self.lib.alll_oracle_priority_mis(n_vars, off, lit, u_ids, len(u_ids), seed, rnd, scratch, out)
        return out[:n].copy()

    def resample(self, off, lit, s_ids, seed, rnd, vars_) -> int:
        s_ids = np.ascontiguousarray(s_ids, np.uint32)
        return int(self.lib.alll_oracle_resample(off, lit, s_ids, len(s_ids), seed, rnd, vars_))

    def round(self, n_vars, off, lit, vars_, seed, rnd):
        """One round in place on ``vars_``; returns (U ascending, S in priority order, n_resampled)."""
        m = len(off) - 1
        u = np.empty(max(m, 1), np.uint32)
        s = np.empty(max(m, 1), np.uint32)
        n_s = C.c_uint64(0)
        n_r = C.c_uint64(0)
        n_u = self.lib.alll_oracle_round(n_vars, m, off, lit, vars_, seed, rnd,
                                         u.ctypes.data, s.ctypes.data, C.byref(n_s), C.byref(n_r))
        return u[:n_u].copy(), s[:n_s.value].copy(), int(n_r.value)

    def solve(self, n_vars, off, lit, vars_, seed, max_rounds=1 << 20, trace=False):
        st = _StatsC()
        tu = ts = None
        if trace:
            tu = np.zeros(max_rounds + 1, np.uint64)
            ts = np.zeros(max_rounds + 1, np.uint64)
        rc = self.lib.alll_oracle_solve(n_vars, len(off) - 1, off, lit, vars_, seed, max_rounds, C.byref(st),
                                        tu.ctypes.data if trace else None, ts.ctypes.data if trace else None)
        stats = Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals, rc)
        if trace:
            return stats, tu[: st.n_iterations].copy(), ts[: st.n_iterations].copy()
        return stats

    def solve_greedy(self, n_vars, off, lit, vars_, seed, n_threads=1, max_rounds=1 << 20) -> Stats:
        st = _StatsC()
        rc = self.lib.alll_oracle_solve_greedy(n_vars, len(off) - 1, off, lit, vars_, seed, n_threads,
                                               max_rounds, C.byref(st))
        return Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals, rc)


class Reference:
    """The unmodified reference headers (oracle/_ref/liballl_ref.so)."""

    def __init__(self):
        if not os.path.exists(_REF_SO):
            raise FileNotFoundError(f"{_REF_SO} missing: run `make -C oracle ref` where /root/reference exists")
        L = self.lib = C.CDLL(_REF_SO)
        L.ref_create.restype = C.c_void_p
        L.ref_create.argtypes = [C.c_uint32, C.c_uint64, _u64p, _u32p, C.c_int]
        L.ref_destroy.argtypes = [C.c_void_p]
        L.ref_set_assignment.argtypes = [C.c_void_p, _u8p]
        L.ref_get_assignment.argtypes = [C.c_void_p, _u8p]
        L.ref_sweep.restype = C.c_uint64
        L.ref_sweep.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_verify.restype = C.c_int
        L.ref_verify.argtypes = [C.c_void_p]
        L.ref_dependent.restype = C.c_int
        L.ref_dependent.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32]
        L.ref_greedy_mis.restype = C.c_uint64
        L.ref_greedy_mis.argtypes = [C.c_void_p, _u32p]
        L.ref_solve.restype = C.c_double
        L.ref_solve.argtypes = [C.c_void_p, _u64p]
        L.ref_rerandomize.argtypes = [C.c_void_p]
        L.ref_time_verify.restype = C.c_double
        L.ref_time_verify.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.ref_num_procs.restype = C.c_int
        L.ref_cnf_header_read.restype = C.c_int
        L.ref_cnf_header_read.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_cnf_data_read.restype = C.c_int
        L.ref_cnf_data_read.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, _i32p, _i32p]
        L.ref_cnf_evaluate.restype = C.c_int
        L.ref_cnf_evaluate.argtypes = [C.c_int, C.c_int, C.c_int, _i32p, _i32p, _u8p]

    def num_procs(self) -> int:
        return int(self.lib.ref_num_procs())

    def instance(self, n_vars, off, lit, n_threads=1) -> "RefInstance":
        return RefInstance(self, n_vars, off, lit, n_threads)

    def cnf_read(self, path: str):
        """cnf_header_read + cnf_data_read; returns (v_num, c_num, l_num, l_c_num, l_val) or None on header error."""
        v, c, l = C.c_int(0), C.c_int(0), C.c_int(0)
        if self.lib.ref_cnf_header_read(path.encode(), C.byref(v), C.byref(c), C.byref(l)):
            return None
        # slack: the reference can write one clause count past c_num
        l_c_num = np.full(c.value + 8, -12345, np.int32)
        l_val = np.zeros(l.value + 8, np.int32)
        self.lib.ref_cnf_data_read(path.encode(), v.value, c.value, l.value, l_c_num, l_val)
        return v.value, c.value, l.value, l_c_num, l_val

    def cnf_evaluate(self, v_num, l_c_num, l_val, v_val) -> bool:
        l_c_num = np.ascontiguousarray(l_c_num, np.int32)
        l_val = np.ascontiguousarray(l_val, np.int32)
        return bool(self.lib.ref_cnf_evaluate(v_num, len(l_c_num), len(l_val), l_c_num, l_val,
                                              np.ascontiguousarray(v_val, np.uint8)))


class RefInstance:
    def __init__(self, ref: Reference, n_vars, off, lit, n_threads):
        self.ref = ref
        self.n_vars = int(n_vars)
        self.m = len(off) - 1
        self.n_threads = n_threads
        self.h = ref.lib.ref_create(self.n_vars, self.m, off, lit, n_threads)

    def close(self):
        if self.h:
            self.ref.lib.ref_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_assignment(self, vars_):
        self.ref.lib.ref_set_assignment(self.h, np.ascontiguousarray(vars_, np.uint8))

    def get_assignment(self) -> np.ndarray:
        out = np.zeros(max(self.n_vars, 1), np.uint8)
        self.ref.lib.ref_get_assignment(self.h, out)
        return out[: self.n_vars]

    def sweep(self) -> np.ndarray:
        out = np.empty(max(self.m, 1), np.uint32)
        n = self.ref.lib.ref_sweep(self.h, out.ctypes.data)
        return out[:n].copy()

    def verify(self) -> bool:
        return bool(self.ref.lib.ref_verify(self.h))

    def dependent(self, c1, c2) -> bool:
        return bool(self.ref.lib.ref_dependent(self.h, c1, c2))

    def greedy_mis(self) -> np.ndarray:
        out = np.empty(max(self.m, 1), np.uint32)
        n = self.ref.lib.ref_greedy_mis(self.h, out)
        return out[:n].copy()

    def rerandomize(self):
        self.ref.lib.ref_rerandomize(self.h)

    def solve(self) -> Stats:
        st = np.zeros(4, np.uint64)
        secs = self.ref.lib.ref_solve(self.h, st)
        return Stats(int(st[0]), int(st[1]), int(st[2]), seconds=float(secs))

    def time_verify(self, reps: int):
        ok = C.c_int(0)
        secs = self.ref.lib.ref_time_verify(self.h, reps, C.byref(ok))
        return float(secs), bool(ok.value)
