"""ctypes binding of the C ABI (include/alll_b200.h) -- the only way Python reaches the kernels.

There is no CPU fallback: if ``liballl_b200.so`` is missing, or no B200-class device is
present, every entry point raises.  (The CPU oracle lives in ``oracle/`` and is test
infrastructure; this package never imports it.)
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALLL_B200_LIB") or os.path.join(_HERE, "liballl_b200.so")   # env override: tuning builds

OK, MAX_ROUNDS, EMPTY_CLAUSE, BAD_ARG, CUDA_ERROR, NCCL_ERROR, NO_INSTANCE, CAPACITY, PREEMPTED = range(9)
STATUS_NAMES = ["OK", "MAX_ROUNDS", "EMPTY_CLAUSE", "BAD_ARG", "CUDA_ERROR", "NCCL_ERROR", "NO_INSTANCE", "CAPACITY", "PREEMPTED"]
FLAG_NO_BUCKETING = 1
FLAG_INCREMENTAL = 4
FLAG_FORCE_CSR = 8
FLAG_P2P_PERSISTENT = 32    # alll_solve_p2p: one persistent kernel per rank (every rank needs its own GPU)
FLAG_HOST_ROUND_LOOP = 16   # alll_solve: one kernel per phase driven by the host instead of the persistent solve kernel
FLAG_FORCE_SHARDING = 64    # alll_multi_*: shard small instances too (tests)
FLAG_NO_PACKING = 128       # sweep streams the plain planes 0..4 instead of the packed eager planes (comparison / tests)

#: every symbol include/alll_b200.h declares (tests check the library exports exactly these)
SYMBOLS = [
    "alll_abi_version", "alll_create", "alll_destroy", "alll_last_error", "alll_device_count", "alll_host_alloc", "alll_host_free",
    "alll_upload_fixedk", "alll_upload_fixedk_device", "alll_upload_csr",
    "alll_set_assignment", "alll_get_assignment", "alll_randomize",
    "alll_eval", "alll_verify", "alll_round", "alll_solve",
    "alll_time_sweep", "alll_launch_count", "alll_layout_info", "alll_sweep_info", "alll_upload_info", "alll_upload_fixedk_streamed",
    "alll_multi_upload_fixedk_streamed",
    "alll_set_id_base", "alll_shard_sweep", "alll_shard_round", "alll_get_stats", "alll_reset_stats",
    "alll_batch_upload", "alll_batch_solve",
    "alll_p2p_create", "alll_p2p_connect", "alll_solve_p2p",
    "alll_upload_generator", "alll_upload_builtin_generator", "alll_builtin_generator_clause",
    "alll_flag_create", "alll_flag_open", "alll_flag_reset", "alll_flag_read", "alll_batch_set_job_base",
    "alll_multi_create", "alll_multi_destroy", "alll_multi_last_error", "alll_multi_upload_fixedk", "alll_multi_upload_csr",
    "alll_multi_set_assignment", "alll_multi_get_assignment", "alll_multi_randomize", "alll_multi_verify", "alll_multi_solve",
    "alll_multi_info", "alll_multi_device_handle", "alll_multi_batch_upload", "alll_multi_batch_solve",
]

GEN_UNIFORM, GEN_BOUNDED = 0, 1


class AlllError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status}: {message}")
        self.status = status


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("sweep_smem_bytes", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32)]


class BatchStatsC(C.Structure):
    _fields_ = [("n_iterations", C.c_uint64), ("n_resamples", C.c_uint64), ("sum_mis_size", C.c_uint64),
                ("status", C.c_int32), ("reserved", C.c_int32)]


class StatsC(C.Structure):
    _fields_ = [("n_iterations", C.c_uint64), ("n_resamples", C.c_uint64), ("avg_mis_size", C.c_uint64),
                ("sum_mis_size", C.c_uint64), ("n_clause_evals", C.c_uint64), ("n_luby_steps", C.c_uint64),
                ("n_kernel_launches", C.c_uint64), ("solve_ms", C.c_double), ("sweep_ms", C.c_double),
                ("status", C.c_int32), ("reserved", C.c_int32), ("between_sweeps_ms", C.c_double),
                ("n_incremental_rounds", C.c_uint64)]


@dataclass
class Stats:
    """Mirror of ``Statistics`` (SATInstance.h:25-32) plus device counters."""
    n_iterations: int
    n_resamples: int
    avg_mis_size: int
    sum_mis_size: int
    n_clause_evals: int
    n_luby_steps: int
    n_kernel_launches: int
    solve_ms: float
    sweep_ms: float
    status: int
    between_sweeps_ms: float = 0.0
    n_incremental_rounds: int = 0


_lib = None


def load() -> C.CDLL:
    """Load the C-ABI library; raises if it has not been built (``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing: build it with `make -C alllsatisfiabilitysolver_b200/csrc` "
            "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u64, u32 = C.c_void_p, C.c_uint64, C.c_uint32
    L.alll_abi_version.restype = C.c_int
    L.alll_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.alll_destroy.argtypes = [vp]
    L.alll_last_error.restype = C.c_char_p
    L.alll_last_error.argtypes = [vp]
    L.alll_upload_fixedk.argtypes = [vp, u64, u64, u32, vp]
    L.alll_upload_fixedk_device.argtypes = [vp, u64, u64, u32, vp]
    L.alll_upload_csr.argtypes = [vp, u64, u64, vp, vp]
    L.alll_upload_generator.argtypes = [vp, u64, u64, u32, vp, vp, u64]
    L.alll_upload_builtin_generator.argtypes = [vp, u32, u64, u64, u32, u64, u32, u64]
    L.alll_builtin_generator_clause.argtypes = [u32, u64, u64, u32, u64, u32, u64, vp]
    L.alll_flag_create.argtypes = [vp, vp]
    L.alll_flag_open.argtypes = [vp, vp]
    L.alll_flag_reset.argtypes = [vp]
    L.alll_flag_read.argtypes = [vp, C.POINTER(C.c_int64)]
    L.alll_batch_set_job_base.argtypes = [vp, u32]
    L.alll_set_assignment.argtypes = [vp, vp]
    L.alll_get_assignment.argtypes = [vp, vp]
    L.alll_randomize.argtypes = [vp, u64]
    L.alll_eval.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.alll_verify.argtypes = [vp, C.POINTER(C.c_int)]
    L.alll_round.argtypes = [vp, u64, u32, vp, u64, C.POINTER(u64), vp, u64, C.POINTER(u64), C.POINTER(u64)]
    L.alll_solve.argtypes = [vp, u64, u64, C.POINTER(StatsC)]
    L.alll_time_sweep.argtypes = [vp, u32, C.POINTER(C.c_double), C.POINTER(u64)]
    L.alll_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.alll_layout_info.argtypes = [vp, C.POINTER(u64)]
    L.alll_sweep_info.argtypes = [vp, C.POINTER(u64)]
    L.alll_upload_info.argtypes = [vp, C.POINTER(u64)]
    L.alll_set_id_base.argtypes = [vp, u64]
    L.alll_shard_sweep.argtypes = [vp, vp, u64, C.POINTER(u64)]
    L.alll_shard_round.argtypes = [vp, vp, C.POINTER(u64), u32, u64, u64, u32, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)]
    L.alll_get_stats.argtypes = [vp, C.POINTER(StatsC)]
    L.alll_reset_stats.argtypes = [vp]
    L.alll_p2p_create.argtypes = [vp, u32, u32, u64, vp]
    L.alll_p2p_connect.argtypes = [vp, vp]
    L.alll_solve_p2p.argtypes = [vp, u64, u64, u64, u32, C.POINTER(StatsC)]
    L.alll_batch_upload.argtypes = [vp, u32, u64, u32, vp, vp]
    L.alll_batch_solve.argtypes = [vp, u32, vp, u64, C.c_int, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    L.alll_device_count.argtypes = [C.POINTER(C.c_int32)]
    L.alll_host_alloc.argtypes = [u64, C.POINTER(vp)]
    L.alll_host_free.argtypes = [vp]
    L.alll_multi_create.argtypes = [vp, u32, C.POINTER(Config), C.POINTER(vp)]
    L.alll_multi_destroy.argtypes = [vp]
    L.alll_multi_last_error.restype = C.c_char_p
    L.alll_multi_last_error.argtypes = [vp]
    L.alll_multi_upload_fixedk.argtypes = [vp, u64, u64, u32, vp]
    L.alll_multi_upload_csr.argtypes = [vp, u64, u64, vp, vp]
    L.alll_multi_set_assignment.argtypes = [vp, vp]
    L.alll_multi_get_assignment.argtypes = [vp, vp]
    L.alll_multi_randomize.argtypes = [vp, u64]
    L.alll_multi_verify.argtypes = [vp, C.POINTER(C.c_int)]
    L.alll_multi_solve.argtypes = [vp, u64, u64, C.POINTER(StatsC)]
    L.alll_multi_info.argtypes = [vp, C.POINTER(u64)]
    L.alll_multi_device_handle.argtypes = [vp, u32, C.POINTER(vp)]
    L.alll_multi_batch_upload.argtypes = [vp, u32, u64, u32, vp, vp]
    L.alll_multi_batch_solve.argtypes = [vp, u32, vp, u64, C.c_int, vp, vp, C.POINTER(C.c_int32), C.POINTER(C.c_double)]
    for name in SYMBOLS:
        getattr(L, name)            # raises AttributeError if the library misses a declared entry point
    _lib = L
    return L


class Solver:
    """Thin owner of one ``alll_handle`` (one CUDA device, one stream)."""

    def __init__(self, device: int = -1, sweep_smem_bytes: int = 0, flags: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        cfg = Config(device, sweep_smem_bytes, flags, 0)
        rc = self.lib.alll_create(C.byref(cfg), C.byref(self.h))
        if rc != OK:
            raise AlllError(rc, self.lib.alll_last_error(None).decode())
        self.n_vars = 0
        self.m = 0
        self._owned = True

    @classmethod
    def borrowed(cls, handle: int, n_vars: int, m: int) -> "Solver":
        """View of an ``alll_handle`` owned by someone else (a device slot of a ``MultiSolver``); ``close`` is a no-op."""
        self = cls.__new__(cls)
        self.lib = load()
        self.h = C.c_void_p(handle)
        self.n_vars, self.m, self._owned = n_vars, m, False
        return self

    # -- plumbing ---------------------------------------------------------------------
    def _check(self, rc: int, allow=(OK,)):
        if rc not in allow:
            raise AlllError(rc, self.lib.alll_last_error(self.h).decode())
        return rc

    def close(self):
        if self.h and self._owned:
            self.lib.alll_destroy(self.h)
        self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- upload -------------------------------------------------------------------------
    def upload_fixedk(self, n_vars: int, lits: np.ndarray):
        lits = np.ascontiguousarray(lits, np.uint32)
        m, k = lits.shape
        self._check(self.lib.alll_upload_fixedk(self.h, n_vars, m, k, lits.ctypes.data))
        self.n_vars, self.m = n_vars, m

    def upload_fixedk_device(self, n_vars: int, m: int, k: int, device_ptr: int):
        """``device_ptr``: raw device address of a row-major [m][k] uint32 buffer (e.g. ``tensor.data_ptr()``)."""
        self._check(self.lib.alll_upload_fixedk_device(self.h, n_vars, m, k, C.c_void_p(device_ptr)))
        self.n_vars, self.m = n_vars, m

    def upload_csr(self, n_vars: int, off: np.ndarray, lit: np.ndarray):
        off = np.ascontiguousarray(off, np.uint64)
        lit = np.ascontiguousarray(lit, np.uint32)
        m = len(off) - 1
        self._check(self.lib.alll_upload_csr(self.h, n_vars, m, off.ctypes.data, lit.ctypes.data if len(lit) else None))
        self.n_vars, self.m = n_vars, m

    def upload_builtin_generator(self, kind: int, n_vars: int, m: int, k: int, seed: int, d: int = 0, cap_records: int = 0):
        """Enumerated clauses (SATInstance.h:70-153): clause i is a pure function of i, nothing is stored."""
        self._check(self.lib.alll_upload_builtin_generator(self.h, kind, n_vars, m, k, seed, d, cap_records))
        self.n_vars, self.m = n_vars, m

    def upload_generator(self, n_vars: int, m: int, k: int, launch_fn: int, user: int = 0, cap_records: int = 0):
        """``launch_fn``: address of an ``alll_gen_launch_fn`` exported by a user-built CUDA library."""
        self._check(self.lib.alll_upload_generator(self.h, n_vars, m, k, launch_fn, user, cap_records))
        self.n_vars, self.m = n_vars, m

    # -- assignment ---------------------------------------------------------------------
    def set_assignment(self, bools: np.ndarray):
        bools = np.ascontiguousarray(bools, np.uint8)
        assert bools.shape == (self.n_vars,)
        self._check(self.lib.alll_set_assignment(self.h, bools.ctypes.data))

    def get_assignment(self, out: np.ndarray | None = None) -> np.ndarray:
        """1 byte per variable, like the reference's ``var_arr->vars``.  ``out``: a caller-owned uint8 buffer of n_vars
        bytes to fill (the reference's array is long-lived too; a fresh 10 MB array costs more in page faults than the copy)."""
        if out is None:
            out = np.empty(self.n_vars, np.uint8)
        elif out.dtype != np.uint8 or out.size != self.n_vars or not out.flags["C_CONTIGUOUS"]:
            raise ValueError("out must be a contiguous uint8 array of n_vars elements")
        self._check(self.lib.alll_get_assignment(self.h, out.ctypes.data))
        return out

    def randomize(self, seed: int):
        self._check(self.lib.alll_randomize(self.h, seed))

    # -- hot path -----------------------------------------------------------------------
    def eval(self, want_ids: bool = True):
        n = C.c_uint64(0)
        if not want_ids:
            self._check(self.lib.alll_eval(self.h, None, 0, C.byref(n)))
            return int(n.value), None
        buf = np.empty(max(self.m, 1), np.uint32)
        self._check(self.lib.alll_eval(self.h, buf.ctypes.data, len(buf), C.byref(n)))
        return int(n.value), buf[: n.value].copy()

    def verify(self) -> bool:
        v = C.c_int(0)
        self._check(self.lib.alll_verify(self.h, C.byref(v)))
        return bool(v.value)

    def round(self, seed: int, rnd: int):
        """One Moser-Tardos round; returns (U ids, S ids, n_resampled), ids in unspecified order."""
        u = np.empty(max(self.m, 1), np.uint32)
        s = np.empty(max(self.m, 1), np.uint32)
        n_u, n_s, n_r = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.alll_round(self.h, seed, rnd, u.ctypes.data, len(u), C.byref(n_u),
                                        s.ctypes.data, len(s), C.byref(n_s), C.byref(n_r)))
        return u[: n_u.value].copy(), s[: n_s.value].copy(), int(n_r.value)

    def solve(self, seed: int, max_rounds: int = 1 << 20) -> Stats:
        st = StatsC()
        self._check(self.lib.alll_solve(self.h, seed, max_rounds, C.byref(st)), allow=(OK, MAX_ROUNDS))
        return Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals,
                     st.n_luby_steps, st.n_kernel_launches, st.solve_ms, st.sweep_ms, st.status, st.between_sweeps_ms, st.n_incremental_rounds)

    # -- clause-range sharded mode (device pointers; the all-gather between the two calls is the caller's) -------
    def set_id_base(self, id_base: int):
        self._check(self.lib.alll_set_id_base(self.h, id_base))

    def shard_sweep(self, d_records_ptr: int, cap_records: int) -> int:
        n = C.c_uint64(0)
        self._check(self.lib.alll_shard_sweep(self.h, C.c_void_p(d_records_ptr), cap_records, C.byref(n)))
        return int(n.value)

    def shard_round(self, d_records_ptr: int, counts, block_cap: int, seed: int, rnd: int):
        arr = (C.c_uint64 * len(counts))(*[int(c) for c in counts])
        n_t, n_s, n_r = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
        self._check(self.lib.alll_shard_round(self.h, C.c_void_p(d_records_ptr), arr, len(counts), block_cap, seed, rnd,
                                              C.byref(n_t), C.byref(n_s), C.byref(n_r)))
        return int(n_t.value), int(n_s.value), int(n_r.value)

    def p2p_create(self, world: int, rank: int, cap_records: int) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.alll_p2p_create(self.h, world, rank, cap_records, buf))
        return bytes(buf)

    def p2p_connect(self, handles: list):
        blob = b"".join(handles)
        arr = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        self._check(self.lib.alll_p2p_connect(self.h, arr))

    def solve_p2p(self, seed: int, m_global: int, epoch: int, max_rounds: int = 1 << 19) -> Stats:
        st = StatsC()
        self._check(self.lib.alll_solve_p2p(self.h, seed, max_rounds, m_global, epoch, C.byref(st)), allow=(OK, MAX_ROUNDS))
        return Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals,
                     st.n_luby_steps, st.n_kernel_launches, st.solve_ms, st.sweep_ms, st.status, st.between_sweeps_ms, st.n_incremental_rounds)

    def get_stats(self) -> Stats:
        st = StatsC()
        self._check(self.lib.alll_get_stats(self.h, C.byref(st)))
        return Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals,
                     st.n_luby_steps, st.n_kernel_launches, st.solve_ms, st.sweep_ms, st.status)

    def reset_stats(self):
        self._check(self.lib.alll_reset_stats(self.h))

    # -- batched small instances / seed portfolio -------------------------------------------------------
    def batch_upload(self, n_vars: int, k: int, clause_off: np.ndarray, lits: np.ndarray):
        """``lits``: row-major (total_clauses, k) uint32; ``clause_off``: (n_instances+1,) row offsets."""
        clause_off = np.ascontiguousarray(clause_off, np.uint64)
        lits = np.ascontiguousarray(lits, np.uint32)
        self._check(self.lib.alll_batch_upload(self.h, len(clause_off) - 1, n_vars, k, clause_off.ctypes.data,
                                               lits.ctypes.data if lits.size else None))
        self._batch = (len(clause_off) - 1, n_vars)

    def batch_solve(self, seeds, max_rounds: int = 1 << 20, portfolio: bool = False, want_assignments: bool = True):
        """Returns (stats structured array, assignments [n_jobs, n_vars] or None, winner, device_ms)."""
        seeds = np.ascontiguousarray(seeds, np.uint64)
        n_jobs, n_vars = len(seeds), self._batch[1]
        stats = np.zeros(n_jobs, dtype=np.dtype([("n_iterations", "<u8"), ("n_resamples", "<u8"), ("sum_mis_size", "<u8"),
                                                 ("status", "<i4"), ("reserved", "<i4")]))
        assign = np.zeros((n_jobs, n_vars), np.uint8) if want_assignments else None
        winner, ms = C.c_int32(-1), C.c_double(0.0)
        self._check(self.lib.alll_batch_solve(self.h, n_jobs, seeds.ctypes.data, max_rounds, int(portfolio),      # 2: shared flag
                                              assign.ctypes.data if want_assignments else None, stats.ctypes.data,
                                              C.byref(winner), C.byref(ms)))
        return stats, assign, int(winner.value), float(ms.value)

    # -- multi-GPU portfolio: one first-SAT word for all ranks --------------------------------
    def flag_create(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._check(self.lib.alll_flag_create(self.h, buf))
        return bytes(buf)

    def flag_open(self, handle: bytes):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.lib.alll_flag_open(self.h, buf))

    def flag_reset(self):
        self._check(self.lib.alll_flag_reset(self.h))

    def flag_read(self) -> int:
        v = C.c_int64(-1)
        self._check(self.lib.alll_flag_read(self.h, C.byref(v)))
        return int(v.value)

    def batch_set_job_base(self, base: int):
        self._check(self.lib.alll_batch_set_job_base(self.h, base))

    # -- measurement ----------------------------------------------------------------------
    def time_sweep(self, reps: int):
        ms = C.c_double(0.0)
        n = C.c_uint64(0)
        self._check(self.lib.alll_time_sweep(self.h, reps, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def launch_count(self) -> int:
        n = C.c_uint64(0)
        self._check(self.lib.alll_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def layout_info(self) -> dict:
        info = (C.c_uint64 * 6)()
        self._check(self.lib.alll_layout_info(self.h, info))
        return dict(m=info[0], k=info[1], n_buckets=info[2], m_padded=info[3], literal_bytes=info[4], sweep_smem_bytes=info[5])

    def sweep_info(self) -> dict:
        info = (C.c_uint64 * 4)()
        self._check(self.lib.alll_sweep_info(self.h, info))
        return dict(packed=bool(info[0]), bucket_relative_literals=info[1], streamed_bytes_per_clause=info[2], min_resident=info[3])

    def upload_info(self) -> dict:
        """How the last host-buffer upload crossed the link (packed H2D transport, include/alll_b200.h)."""
        info = (C.c_uint64 * 4)()
        self._check(self.lib.alll_upload_info(self.h, info))
        return dict(link_bytes=info[0], packed_chunks=info[1], raw_chunks=info[2], pack_threads=info[3])


class MultiSolver:
    """Several GPUs behind one call from ONE process (``alll_multi_*``): clause-range sharded solve of one large
    instance (every device uploads only its own range from the host buffer; fused NVLink exchange, peer access instead
    of CUDA IPC), or batched small instances / a seed portfolio spread over the devices.  The same device may be listed
    several times (simulated ranks on one GPU)."""

    def __init__(self, devices, sweep_smem_bytes: int = 0, flags: int = 0):
        self.lib = load()
        self.h = C.c_void_p()
        devs = (C.c_int32 * len(devices))(*[int(d) for d in devices])
        cfg = Config(-1, sweep_smem_bytes, flags, 0)
        rc = self.lib.alll_multi_create(devs, len(devices), C.byref(cfg), C.byref(self.h))
        if rc != OK:
            raise AlllError(rc, self.lib.alll_multi_last_error(None).decode())
        self.n_devices = len(devices)
        self.n_vars = 0
        self.m = 0

    def _check(self, rc: int, allow=(OK,)):
        if rc not in allow:
            raise AlllError(rc, self.lib.alll_multi_last_error(self.h).decode())
        return rc

    def close(self):
        if self.h:
            self.lib.alll_multi_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def upload_fixedk(self, n_vars: int, lits: np.ndarray):
        lits = np.ascontiguousarray(lits, np.uint32)
        m, k = lits.shape
        self._check(self.lib.alll_multi_upload_fixedk(self.h, n_vars, m, k, lits.ctypes.data))
        self.n_vars, self.m = n_vars, m

    def upload_csr(self, n_vars: int, off: np.ndarray, lit: np.ndarray):
        off = np.ascontiguousarray(off, np.uint64)
        lit = np.ascontiguousarray(lit, np.uint32)
        self._check(self.lib.alll_multi_upload_csr(self.h, n_vars, len(off) - 1, off.ctypes.data, lit.ctypes.data if len(lit) else None))
        self.n_vars, self.m = n_vars, len(off) - 1

    def set_assignment(self, bools: np.ndarray):
        bools = np.ascontiguousarray(bools, np.uint8)
        assert bools.shape == (self.n_vars,)
        self._check(self.lib.alll_multi_set_assignment(self.h, bools.ctypes.data))

    def get_assignment(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.n_vars, np.uint8)
        self._check(self.lib.alll_multi_get_assignment(self.h, out.ctypes.data))
        return out

    def randomize(self, seed: int):
        self._check(self.lib.alll_multi_randomize(self.h, seed))

    def verify(self) -> bool:
        v = C.c_int(0)
        self._check(self.lib.alll_multi_verify(self.h, C.byref(v)))
        return bool(v.value)

    def solve(self, seed: int, max_rounds: int = 1 << 19) -> Stats:
        st = StatsC()
        self._check(self.lib.alll_multi_solve(self.h, seed, max_rounds, C.byref(st)), allow=(OK, MAX_ROUNDS))
        return Stats(st.n_iterations, st.n_resamples, st.avg_mis_size, st.sum_mis_size, st.n_clause_evals,
                     st.n_luby_steps, st.n_kernel_launches, st.solve_ms, st.sweep_ms, st.status, st.between_sweeps_ms, st.n_incremental_rounds)

    def device_solver(self, i: int) -> Solver:
        """Device slot ``i`` as a (borrowed) ``Solver``: its clause range only."""
        h = C.c_void_p()
        self._check(self.lib.alll_multi_device_handle(self.h, i, C.byref(h)))
        return Solver.borrowed(h.value, self.n_vars, self.m)

    def info(self) -> dict:
        info = (C.c_uint64 * 4)()
        self._check(self.lib.alll_multi_info(self.h, info))
        return dict(devices_in_use=info[0], sharded=bool(info[1]), widest_range=info[2], cap_records=info[3])

    def batch_upload(self, n_vars: int, k: int, clause_off: np.ndarray, lits: np.ndarray):
        clause_off = np.ascontiguousarray(clause_off, np.uint64)
        lits = np.ascontiguousarray(lits, np.uint32)
        self._check(self.lib.alll_multi_batch_upload(self.h, len(clause_off) - 1, n_vars, k, clause_off.ctypes.data,
                                                     lits.ctypes.data if lits.size else None))
        self._batch = (len(clause_off) - 1, n_vars)

    def batch_solve(self, seeds, max_rounds: int = 1 << 20, portfolio: bool = False, want_assignments: bool = True):
        """Returns (stats structured array, assignments [n_jobs, n_vars] or None, winner, device_ms max over devices)."""
        seeds = np.ascontiguousarray(seeds, np.uint64)
        n_jobs, n_vars = len(seeds), self._batch[1]
        stats = np.zeros(n_jobs, dtype=np.dtype([("n_iterations", "<u8"), ("n_resamples", "<u8"), ("sum_mis_size", "<u8"),
                                                 ("status", "<i4"), ("reserved", "<i4")]))
        assign = np.zeros((n_jobs, n_vars), np.uint8) if want_assignments else None
        winner, ms = C.c_int32(-1), C.c_double(0.0)
        self._check(self.lib.alll_multi_batch_solve(self.h, n_jobs, seeds.ctypes.data, max_rounds, int(portfolio),
                                                    assign.ctypes.data if want_assignments else None, stats.ctypes.data,
                                                    C.byref(winner), C.byref(ms)))
        return stats, assign, int(winner.value), float(ms.value)


def builtin_generator_clauses(kind: int, n_vars: int, m: int, k: int, seed: int, d: int = 0, indices=None) -> np.ndarray:
    """Host evaluation of a built-in generator (no GPU needed): the (len(indices), k) literal matrix."""
    lib = load()
    idx = np.arange(m, dtype=np.uint64) if indices is None else np.asarray(indices, np.uint64)
    out = np.empty((len(idx), k), np.uint32)
    row = np.empty(k, np.uint32)
    for r, i in enumerate(idx):
        rc = lib.alll_builtin_generator_clause(kind, n_vars, m, k, seed, d, int(i), row.ctypes.data)
        if rc != OK:
            raise AlllError(rc, "alll_builtin_generator_clause")
        out[r] = row
    return out
