#!/usr/bin/env python
"""bench.py -- the reference's headline metric on the north-star workload (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl reference]

A *step* is one complete Moser-Tardos solve (random assignment -> all clauses satisfied) of the workload
instance; the metric is clause-evals/sec = m x sweeps / device time of the round loop, with resample
rounds/sec, time-to-SAT and the sweep kernel's HBM roofline alongside.  Workload at N=1: BASELINE config 4,
bounded-degree 8-SAT n=10M m~40M -- the configuration the north-star target is quoted on; its 1.28 GB literal
stream is ~10x the 126 MB L2, so no L2 flush is needed between iterations.

Prints ONE JSON line (rank 0).  ``--impl reference`` times the unmodified reference headers
(oracle/_ref, OpenMP, all host threads) on a bounded sample of the same shape.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from alllsatisfiabilitysolver_b200.instances import CONFIGS, INSTANCE_SEED_BASE  # noqa: E402

METRIC = "clause_evals_per_sec"
UNIT = "clause-evals/s"


def workload_shape(name: str, scale: float):
    cfg = dict(CONFIGS[name])
    cfg["n"] = max(int(cfg["n"] * scale), 1000)
    if cfg["kind"] == "uniform":
        cfg["m"] = max(int(cfg["m"] * scale), 1000)
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe).

    Samples come from NVML (nvidia_ml_py, one query every ~4 ms in a thread) or, without it, from an `nvidia-smi -lms 20`
    child.  The sampler is started before the warm-up so that it is already running when the timed region begins; every
    sample carries its arrival time and stop(t0, t1) keeps the ones that fell inside the region.
    """

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.samples = []          # (t, sm_mhz, sm_max_mhz, [reason names])
        self.proc = None
        self.source = None
        self._stop = threading.Event()
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.bits = [(getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8), "hw_slowdown"),
                         (getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                         (getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                         (getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4), "sw_power_cap")]
            self._sample_nvml()                                  # fails here rather than in the thread
            self.source = "nvml"
            self.t = threading.Thread(target=self._loop_nvml, daemon=True)
            self.t.start()
            return
        except Exception:
            self.samples = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _sample_nvml(self):
        nv = self.nv
        mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
        self.samples.append((time.perf_counter(), mhz, self.max_mhz, [nm for bit, nm in self.bits if mask & bit]))

    def _loop_nvml(self):
        while not self._stop.is_set():
            try:
                self._sample_nvml()
            except Exception:
                return
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 6:
                try:
                    self.samples.append((time.perf_counter(), float(parts[0]), float(parts[1]),
                                         [nm for nm, val in zip(self.NAMES, parts[2:6]) if val.lower().startswith("active")]))
                except ValueError:
                    continue

    def stop(self, t0: float | None = None, t1: float | None = None):
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"]}
        if self.source == "nvidia-smi":
            time.sleep(0.05)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        else:
            self._stop.set()
            self.t.join(timeout=1)
        inside = [s for s in self.samples if t0 is not None and t0 <= s[0] <= t1]
        use = inside if inside else self.samples
        sm = [s[1] for s in use]
        reasons = sorted({r for s in use for r in s[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((s[2] for s in use), default=None),
                "reasons": reasons, "samples": len(sm), "source": self.source,
                "window": "timed region" if inside else "whole run (no sample fell inside the timed region)"}


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the unmodified reference headers on host cores (oracle/_ref)
# ------------------------------------------------------------------------------------------------

SAMPLE_N = 1_000_000      # variables of the bounded sample both arms can solve (see sample_instance)


def sample_instance(shape: dict):
    """The bounded sample of the workload: same (k, d) shape at n = 1M variables at most, numpy generator, fixed seed --
    the EXACT instance the reference arm solves and the `same_instance` record of our arm solves on the GPU.
    (Full cfg4 needs ~5 GB of Clause objects and ~10 minutes per solve in the reference's quadratic greedy independent
    set: 601.6 s measured with 8 threads, profiles/r02_reference_full_cfg4.md.)"""
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat, uniform_ksat

    n = min(shape["n"], SAMPLE_N)
    if shape["kind"] == "bounded":
        lits = bounded_degree_ksat(n, shape["k"], shape["d"], seed=INSTANCE_SEED_BASE)
        desc = f"bounded-degree {shape['k']}-SAT d={shape['d']} n={n} m={lits.shape[0]}"
    else:
        m = int(shape["m"] * n / shape["n"])
        lits = uniform_ksat(n, shape["k"], m, seed=INSTANCE_SEED_BASE)
        desc = f"uniform {shape['k']}-SAT n={n} m={m}"
    return n, lits, desc


def reference_sample(shape: dict, steps: int, warmup: int, budget_s: float = 150.0):
    """Times SATInstance::solve (the -p OpenMP path) on a bounded sample of the workload shape."""
    from oracle.oracle import Reference, have_reference, to_csr

    if not have_reference():
        return None
    ref = Reference()
    cores = ref.num_procs()
    n, lits, desc = sample_instance(shape)
    off, lit = to_csr(lits)
    m = lits.shape[0]
    evals, secs, its = 0, 0.0, []
    t_start = time.time()
    with ref.instance(n, off, lit, cores) as ri:
        done = 0
        for i in range(warmup + steps):
            ri.rerandomize()
            st = ri.solve()
            assert ri.verify(), "reference produced an invalid assignment"
            if i >= warmup:
                evals += m * st.n_iterations
                secs += st.seconds
                its.append(st.n_iterations)
                done += 1
            if time.time() - t_start > budget_s and done >= 1:
                break
    return dict(value=evals / secs, unit=UNIT, cores=cores, kind="reference",
                sample=f"{desc}; {done} full SATInstance::solve runs with n_threads={cores} (unmodified reference headers, "
                       f"-Ofast -fopenmp), mean {np.mean(its):.1f} iterations, {secs / done * 1e3:.1f} ms per solve",
                ms_per_step=secs / done * 1e3, steps=done, n=n, m=m, desc=desc)


_FULL_REF_CHILD = r"""
import sys, time, json
sys.path.insert(0, {root!r})
import numpy as np
from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat, uniform_ksat, INSTANCE_SEED_BASE
from oracle.oracle import Reference, to_csr
shape = {shape!r}
t0 = time.time()
lits = (bounded_degree_ksat(shape["n"], shape["k"], shape["d"], seed=INSTANCE_SEED_BASE) if shape["kind"] == "bounded"
        else uniform_ksat(shape["n"], shape["k"], shape["m"], seed=INSTANCE_SEED_BASE))
off, lit = to_csr(lits)
ref = Reference()
with ref.instance(shape["n"], off, lit, ref.num_procs()) as ri:
    print(json.dumps({{"built_s": time.time() - t0, "m": int(lits.shape[0])}}), flush=True)
    st = ri.solve()
    print(json.dumps({{"solve_s": st.seconds, "n_iterations": st.n_iterations, "verified": bool(ri.verify())}}), flush=True)
"""


def reference_full_instance(shape: dict, timebox_s: float):
    """The reference ONCE on the full workload instance, in a child process that is stopped at the time box (the solve
    is one blocking C++ call).  Returns what happened -- a result, or 'no result in T s' (SURVEY 8d expected minutes)."""
    import tempfile

    if timebox_s <= 0:
        return {"skipped": "time box 0"}
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
        f.write(_FULL_REF_CHILD.format(root=ROOT, shape=dict(shape)))
        path = f.name
    t0 = time.time()
    out = ""
    try:
        proc = subprocess.Popen([sys.executable, path], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, start_new_session=True)
        try:
            out, _ = proc.communicate(timeout=timebox_s)
        except subprocess.TimeoutExpired:
            import signal
            os.killpg(proc.pid, signal.SIGKILL)              # the exact process group we started
            out, _ = proc.communicate()
    finally:
        os.unlink(path)
    info = {}
    for line in (out or "").splitlines():
        try:
            info.update(json.loads(line))
        except Exception:
            pass
    if "solve_s" in info:
        return {"workload": describe(shape, "full"), "result": info, "clause_evals_per_sec": info["m"] * info["n_iterations"] / info["solve_s"]}
    return {"workload": describe(shape, "full"), "result": f"no result in {timebox_s:.0f} s (time box; instance built: {'built_s' in info})",
            "elapsed_s": time.time() - t0,
            "measured_elsewhere": "601.6 s per solve (17 iterations, verified) with 8 threads in the build container: profiles/r02_reference_full_cfg4.md"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    shape = workload_shape(args.workload, args.scale)
    r = reference_sample(shape, args.steps, args.warmup)
    if r is None:
        emit_line({"impl": "reference", "unavailable": "oracle/_ref/liballl_ref.so not built"})
        return 0
    sampled = r["n"] != shape["n"]
    full = None
    if sampled and args.full_ref_timebox > 0:
        full = reference_full_instance(shape, args.full_ref_timebox)
    # config.workload says what THIS arm solved: the bounded sample, not the full instance of our arm's `value`; the
    # like-for-like comparison is our arm's `same_instance` record (the GPU on exactly this sample instance)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": r["steps"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": (f"{args.workload} SHAPE, bounded sample: {r['desc']} (the full instance is n={shape['n']})" if sampled
                                    else describe(shape, args.workload)),
                       "sample": r["sample"], "same_instance_as": "our arm's `same_instance` record"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "full_instance": full,
            "gpu_launches": 0}
    emit_line(line)
    return 0


def describe(shape, name):
    if shape["kind"] == "bounded":
        return f"{name}: bounded-degree random {shape['k']}-SAT, n={shape['n']} vars, every var <= {shape['d']} occurrences"
    return f"{name}: uniform random {shape['k']}-SAT, n={shape['n']} vars, m={shape['m']} clauses"


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------

def independent_check(lits, assignment: np.ndarray) -> bool:
    """north_star: 'an independent checker verifies every returned assignment'.  NOT the GPU's own sweep: the caller's
    literals are turned into signed DIMACS literals and evaluated on the host by the oracle's restatement of the
    reference's cnf_evaluate (cnf_io.cpp:392-484) -- checker use of oracle/, outside every timed region.
    ``lits``: (m, k) torch tensor on the device (int32 view of uint32 2*var+neg) or numpy uint32 matrix."""
    from oracle.oracle import Oracle

    if isinstance(lits, np.ndarray):
        flat = lits.reshape(-1)
        var1 = (flat >> 1).astype(np.int64) + 1
        l_val = np.where(flat & 1, -var1, var1).astype(np.int32)
        m, k = lits.shape
    else:
        import torch

        m, k = int(lits.shape[0]), int(lits.shape[1])
        var1 = (lits >> 1) + 1
        l_val = torch.where((lits & 1) != 0, -var1, var1).to(torch.int32).reshape(-1).cpu().numpy()
        del var1
    return bool(Oracle().check_signed(np.full(m, k, np.int32), l_val, np.ascontiguousarray(assignment, np.uint8)))


def same_instance_record(shape: dict, device: int, cpu: dict | None, max_rounds: int):
    """Like-for-like: the GPU arm on the EXACT instance the reference arm / cpu_baseline solves (sample_instance), device
    timed and end to end from host buffers, each result checked by the independent checker."""
    import torch

    from alllsatisfiabilitysolver_b200 import capi

    n, lits, desc = sample_instance(shape)
    m, k = lits.shape
    host_t = torch.from_numpy(lits.view(np.int32)).pin_memory()
    host = host_t.numpy().view(np.uint32)
    out_t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    out = out_t.numpy()
    ok = True
    with capi.Solver(device=device) as s:
        s.upload_fixedk(n, host)
        dev = []
        for i in range(-3, 10):
            s.randomize(4000 + i)
            st = s.solve(4000 + i, max_rounds)
            if i >= 0:
                dev.append(st)
        ok &= all(x.status == 0 for x in dev) and independent_check(lits, s.get_assignment())
        e2e_ms, e2e_evals = [], 0

        def step(i):
            s.upload_fixedk(n, host)
            s.randomize(5000 + i)
            st = s.solve(5000 + i, max_rounds)
            s.get_assignment(out)
            return st

        step(-1)
        for i in range(5):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            st = step(i)
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            e2e_evals += st.n_clause_evals
        ok &= independent_check(lits, out)
    gpu_value = sum(x.n_clause_evals for x in dev) / (sum(x.solve_ms for x in dev) * 1e-3)
    gpu_e2e = e2e_evals / (sum(e2e_ms) * 1e-3)
    rec = {"workload": desc, "instance": "numpy generator, seed INSTANCE_SEED_BASE: identical literals in both arms",
           "gpu": {"value": gpu_value, "unit": UNIT, "time_to_sat_ms": float(np.mean([x.solve_ms for x in dev])),
                   "sweeps_per_solve": float(np.mean([x.n_iterations for x in dev])), "solves": len(dev)},
           "gpu_e2e": {"value": gpu_e2e, "unit": UNIT, "ms_per_step": float(np.mean(e2e_ms)), "steps": len(e2e_ms),
                       "h2d_bytes_per_step": 4 * k * m, "d2h_bytes_per_step": n,
                       "call": "alll_upload_fixedk(host) + alll_randomize + alll_solve + alll_get_assignment(host)"},
           "all_verified_by_independent_checker": bool(ok)}
    if cpu:
        rec["reference"] = {"value": cpu["value"], "unit": UNIT, "cores": cpu["cores"], "ms_per_solve": cpu.get("ms_per_step"),
                            "note": "measured in this same run on this box's host cores (cpu_baseline)"}
        rec["ratio_device_timed"] = gpu_value / cpu["value"]
        rec["ratio_e2e"] = gpu_e2e / cpu["value"]
        if cpu.get("ms_per_step"):
            rec["time_to_sat_ratio"] = cpu["ms_per_step"] / rec["gpu"]["time_to_sat_ms"]
    return rec


_FULL_AFFINITY = None      # this process's CPU set before bind_to_gpu_numa_node narrowed it (the CPU baseline gets it back)


def bind_to_gpu_numa_node(index):
    """Pins this process to the host CPUs NVML names as local to GPU `index`, before any page-locked buffer exists: the
    end-to-end steps copy 1.28 GB per step and rank from pinned host memory, and with one rank per GPU all of them pull at
    once -- buffers first-touched on the far socket would cross the inter-socket link.  Returns a small record for the line
    (None when NVML or the affinity call is unavailable: nothing changes then)."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        n_cpu = os.cpu_count() or 1
        mask = nv.nvmlDeviceGetCpuAffinity(h, (n_cpu + 63) // 64)
        global _FULL_AFFINITY
        _FULL_AFFINITY = set(os.sched_getaffinity(0))
        cpus = {i for i in range(n_cpu) if (int(mask[i // 64]) >> (i % 64)) & 1} & _FULL_AFFINITY
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return {"cpus": len(cpus), "first": min(cpus), "last": max(cpus), "source": "nvmlDeviceGetCpuAffinity"}
    except Exception as e:                                  # noqa: BLE001 -- a missing NVML must not cost the bench line
        return {"error": repr(e)[:80]}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat_torch, uniform_ksat_torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) goes to stdout too
    if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
        os.environ["NCCL_DEBUG"] = "WARN"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the solver has no CPU fallback")
    host_affinity = None if os.environ.get("ALLL_BENCH_NO_AFFINITY") else bind_to_gpu_numa_node(local_rank)
    torch.cuda.set_device(local_rank)
    if world > 1:
        # N ranks copy their replicas from ONE host at the same time: N x the pack threads would oversubscribe its cores
        os.environ.setdefault("ALLL_H2D_PACK", "0")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    host_group = dist.new_group(backend="gloo") if world > 1 else None     # barriers that keep the waiting ranks' GPUs idle

    shape = workload_shape(args.workload, args.scale)
    k, n = shape["k"], shape["n"]
    # weak scaling over the natural shard: one independent solver per GPU, no exchange.  Every GPU gets its own copy of the
    # SAME instance and the same solve seeds, so the per-GPU work is exactly fixed as N grows (random instances of one shape
    # differ in the number of sweeps their solves need -- 15.6 vs 16.6 per solve between two seeds -- and MAX over ranks then
    # reports that spread instead of the system); the copies must also end bit-identical, which the line checks.
    inst_seed = INSTANCE_SEED_BASE + int(args.workload[3:])
    if shape["kind"] == "bounded":
        lits_t = bounded_degree_ksat_torch(n, k, shape["d"], inst_seed)
    else:
        lits_t = uniform_ksat_torch(n, k, shape["m"], inst_seed)
    m = int(lits_t.shape[0])
    torch.cuda.synchronize()

    solver = capi.Solver(device=local_rank)
    solver.upload_fixedk_device(n, m, k, lits_t.data_ptr())
    layout = solver.layout_info()
    max_rounds = args.max_rounds

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step(i):
        solver.randomize(1000 + i)
        return solver.solve(1000 + i, max_rounds)

    clocks = ClockSampler(local_rank) if rank == 0 else None   # already sampling when the timed region starts
    for i in range(args.warmup):
        one_step(-1 - i)

    # ---- timed region: exactly K steps, inputs resident in HBM ----
    barrier()
    launches0 = solver.launch_count()
    t0 = time.perf_counter()
    stats = [one_step(i) for i in range(args.steps)]
    barrier()
    t1 = time.perf_counter()
    wall_ms = (t1 - t0) * 1e3
    launches = solver.launch_count() - launches0
    clk = clocks.stop(t0, t1) if clocks else None

    dev_ms = sum(s.solve_ms for s in stats)                 # CUDA events on the solver's stream
    evals = sum(s.n_clause_evals for s in stats)
    sweeps = sum(s.n_iterations for s in stats)
    sweep_ms = sum(s.sweep_ms for s in stats)
    rounds = sum(s.n_iterations - 1 for s in stats)
    all_sat = all(s.status == 0 for s in stats)
    # the last timed step's assignment, checked OUTSIDE the timed region by the independent signed-literal checker
    # (cnf_io.cpp:392-484 semantics, restated in oracle/) -- not by the GPU's own sweep
    verified = bool(all_sat and independent_check(lits_t, solver.get_assignment()))
    verified_gpu_sweep = bool(solver.verify())

    # ---- same workload with incremental re-evaluation (opt-in mode; bit-identical results, fewer clause evaluations) ----
    incremental = None
    if not args.no_extras and world == 1:
        inc = capi.Solver(device=local_rank, flags=capi.FLAG_INCREMENTAL)
        inc.upload_fixedk_device(n, m, k, lits_t.data_ptr())
        res = []
        for i in range(-2, args.steps):
            inc.randomize(1000 + max(i, 0))
            st_i = inc.solve(1000 + max(i, 0), max_rounds)
            if i >= 0:
                res.append(st_i)
        same = all((a.n_iterations, a.n_resamples, a.sum_mis_size) == (b.n_iterations, b.n_resamples, b.sum_mis_size)
                   for a, b in zip(res, stats))
        incremental = {"time_to_sat_ms": float(np.mean([r.solve_ms for r in res])),
                       "incremental_rounds_per_solve": float(np.mean([r.n_incremental_rounds for r in res])),
                       "clauses_evaluated_per_solve": float(np.mean([r.n_clause_evals for r in res])),
                       "same_statistics_as_full_sweep_mode": bool(same), "verified": bool(inc.verify()),
                       "note": "ALLL_FLAG_INCREMENTAL: occurrence lists of the resampled variables replace the full sweep "
                               "once few variables are resampled (SURVEY 8f-3); not used for `value`"}
        inc.close()

    # ---- e2e: the reference-facing call with HOST buffers (upload + solve + assignment read-back) ----
    lits_host = torch.empty(lits_t.shape, dtype=lits_t.dtype, pin_memory=True)
    lits_host.copy_(lits_t)
    torch.cuda.synchronize()
    host_np = lits_host.numpy().view(np.uint32)
    e2e_solver = capi.Solver(device=local_rank)
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    e2e_evals = 0
    e2e_out_t = torch.empty(n, dtype=torch.uint8, pin_memory=True)   # the caller's assignment array (the reference's var_arr->vars):
    e2e_out = e2e_out_t.numpy()                                      # long-lived and page-locked, like the literal buffer

    def e2e_step(i):
        e2e_solver.upload_fixedk(n, host_np)
        e2e_solver.randomize(2000 + i)
        st = e2e_solver.solve(2000 + i, max_rounds)
        e2e_solver.get_assignment(e2e_out)
        return st

    for i in range(min(3, max(1, args.warmup))):     # (the first calls grow the pooled buffers and the page-locked ring)
        e2e_step(-1 - i)
    barrier()
    e2e_step_ms = []
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        ts = time.perf_counter()
        e2e_evals += e2e_step(i).n_clause_evals
        e2e_step_ms.append(round((time.perf_counter() - ts) * 1e3, 3))
    barrier()
    e2e_s = time.perf_counter() - t0
    up = e2e_solver.upload_info()                    # how the last step's literals crossed the link (packed H2D transport)
    e2e_solver.close()

    # ---- other BASELINE configs, briefly (parity-test cases, not bench lines): cfg2 single solve, cfg5 batch ----
    extras = None
    if world == 1 and not args.no_extras and args.workload == "cfg4" and args.scale == 1.0:
        try:
            extras = other_workloads(local_rank)
        except Exception as e:                       # never lose the main line to an extra
            extras = {"error": repr(e)}

    # ---- clause-range sharded solve of the SAME workload over all ranks (strong scaling, NCCL all-gather) ----
    sharded = None
    if world > 1 and not args.no_sharded:
        try:
            sharded = sharded_solves(args, shape, rank, world, local_rank, host_group)
        except Exception as e:                       # e.g. CUDA IPC not permitted on this box: report, do not die
            sharded = {"error": repr(e)}

    # ---- BASELINE config 5 over all ranks: seed portfolio with ONE first-SAT word for all GPUs, batched instances ----
    cfg5_multi = None
    if world > 1 and not args.no_sharded:
        try:
            cfg5_multi = cfg5_over_ranks(rank, world, local_rank)
        except Exception as e:
            cfg5_multi = {"error": repr(e)}

    # ---- reduce over ranks: MAX time, SUM work ----
    red = torch.tensor([dev_ms, wall_ms, e2e_s], dtype=torch.float64, device="cuda")
    tot = torch.tensor([evals, sweeps, rounds, e2e_evals, launches, int(all_sat and verified)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max, e2e_s_max = [float(x) for x in red.tolist()]
    evals_all, sweeps_all, rounds_all, e2e_evals_all, launches_all, n_ok = [float(x) for x in tot.tolist()]
    per_rank = None
    if world > 1:
        mine = torch.tensor([dev_ms / args.steps, sweeps / args.steps, sweep_ms / max(sweeps, 1)], dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        # the copies solve the same instance with the same seeds: same statistics and the same final assignment on every GPU
        import zlib
        sig = torch.tensor([zlib.crc32(solver.get_assignment().tobytes()), int(sum(x.n_resamples for x in stats)) & 0x7FFFFFFF],
                           dtype=torch.int64, device="cuda")
        sigs = [torch.zeros_like(sig) for _ in range(world)]
        dist.all_gather(sigs, sig)
        per_rank = {"ms_per_step": [round(float(x[0]), 4) for x in allr], "sweeps_per_solve": [round(float(x[1]), 2) for x in allr],
                    "sweep_phase_ms": [round(float(x[2]), 4) for x in allr],
                    "copies_bit_identical": bool(all(torch.equal(x, sigs[0]) for x in sigs)),
                    "note": "every rank solves its own copy of the same instance with the same seeds (identical work per GPU)"}

    # ---- the sweep as a kernel of its own, back to back on the same resident data (CUDA events, alll_time_sweep):
    # the cross-check for the roofline figure below, whose kernel runs every sweep of a solve in one launch ----
    solver.randomize(999)
    standalone_ms, standalone_viol = solver.time_sweep(20)

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = 4 * k * m + n // 8               # 4k bytes per clause-eval + the packed assignment once
        # alll_solve is ONE cooperative launch (sweep -> independent set + resample -> sweep ... on the device), so the
        # dominant kernel's launch is the whole solve: algorithmic bytes per launch = sweeps x (4km + n/8), duration =
        # the CUDA-event time of that launch.  The independent-set phases and grid barriers between the sweeps are
        # inside this duration, i.e. the figure is a lower bound on the bandwidth of the sweep phase itself, which the
        # kernel also times (%globaltimer read by block 0 around each sweep and its grid barrier).
        persistent = all(s.n_kernel_launches <= 3 for s in stats)
        launch_ms = dev_ms / args.steps
        alg_bytes_launch = alg_bytes * sweeps / args.steps
        achieved = alg_bytes_launch / (launch_ms * 1e-3) / 1e9
        sweep_avg_ms = sweep_ms / max(sweeps, 1)
        sweep_phase = alg_bytes / (sweep_avg_ms * 1e-3) / 1e9 if sweep_avg_ms > 0 else 0.0
        standalone = alg_bytes / (standalone_ms * 1e-3) / 1e9 if standalone_ms > 0 else 0.0
        traffic, traffic_source = None, None
        tp = os.path.join(ROOT, "profiles", "sweep_traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                # dram bytes of ONE solve launch from the ncu capture named in the file (that capture's solve ran
                # `sweeps` sweeps); falls back to the per-sweep figure of the standalone kernel x sweeps of this run
                if persistent and tj.get(args.workload + "_solve_launch"):
                    traffic = tj[args.workload + "_solve_launch"] / tj[args.workload + "_solve_launch_sweeps"] * (sweeps / args.steps)
                    traffic_source = (f"NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of one solve_persistent_kernel launch "
                                      f"({tj[args.workload + '_solve_launch_sweeps']} sweeps) from the committed ncu --set full capture "
                                      f"{tj.get('source', 'profiles/sweep_traffic.json')}, rescaled to this run's {sweeps / args.steps:.1f} sweeps per launch")
                elif tj.get(args.workload):
                    traffic = tj[args.workload] * (sweeps / args.steps if persistent else 1)
                    traffic_source = "NOT measured in this run: per-sweep DRAM bytes of the committed ncu capture (profiles/sweep_traffic.json)"
            except Exception:
                traffic = None
        cpu, same = None, None
        if world == 1 and not args.no_cpu_baseline:
            if _FULL_AFFINITY:                                  # the reference runs on ALL host cores, as in its own arm
                os.sched_setaffinity(0, _FULL_AFFINITY)
            cpu_full = reference_sample(shape, steps=3, warmup=1, budget_s=60.0)
            if cpu_full:
                cpu = {kk: cpu_full[kk] for kk in ("value", "unit", "cores", "kind", "sample")}
            if args.scale == 1.0 and not args.no_extras:
                try:
                    same = same_instance_record(shape, local_rank, cpu_full, max_rounds)
                except Exception as e:               # never lose the main line to an extra
                    same = {"error": repr(e)}
        line = {
            "metric": METRIC, "value": evals_all / (dev_ms_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": describe(shape, args.workload), "m_clauses_per_gpu": m, "k": k,
                       "step": "one full solve: random assignment -> verified satisfying assignment",
                       "parallelism": f"{world} x independent solver, one per GPU, each on its own copy of the instance (no exchange)" if world > 1 else "single GPU",
                       "l2": "literal stream (4*k*m bytes) is larger than L2; no flush needed" if 4 * k * m > 126e6 else
                             "literal stream fits L2 (126 MB): sweeps after the first are L2-resident",
                       "layout": layout},
            "host_affinity": host_affinity,
            "per_rank": per_rank,
            "time_to_sat_ms": dev_ms_max / args.steps,
            "rounds_per_sec": rounds_all / (dev_ms_max * 1e-3),
            "sweeps_per_solve": sweeps / args.steps,
            "all_runs_sat_and_verified": bool(n_ok == world),
            "verification": {"checker": "independent: signed DIMACS literals evaluated on the host (oracle restatement of the reference's "
                                        "cnf_evaluate, cnf_io.cpp:392-484), last timed step's assignment of every rank, outside the timed region",
                             "gpu_sweep_agrees": verified_gpu_sweep},
            "same_instance": same,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "roofline": {"bound": "hbm", "kernel": "solve_persistent_kernel" if persistent else "sweep_planes_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_source, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes_launch, "avg_launch_ms": launch_ms,
                         "launch": "one launch = one whole solve (all sweeps + independent-set phases); CUDA events on the solver's stream",
                         "frac_of_nominal_8TBps": achieved / 8000.0,
                         "sweep_phase": {"avg_ms": sweep_avg_ms, "achieved": sweep_phase, "frac": sweep_phase / peak,
                                         "frac_of_nominal_8TBps": sweep_phase / 8000.0, "algorithmic_bytes": alg_bytes,
                                         "timer": "%globaltimer, block 0, around each sweep phase + its grid barrier, inside the timed steps"},
                         "standalone_sweep_kernel": {"avg_launch_ms": standalone_ms, "achieved": standalone, "frac": standalone / peak,
                                                     "violated_per_sweep": standalone_viol,
                                                     "timer": "CUDA events around 20 back-to-back sweep_planes_kernel launches from a random "
                                                              "assignment (the first sweep of a solve: most violated clauses, most record writes)"}},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_evals_all / e2e_s_max, "unit": UNIT, "h2d_bytes_per_step": 4 * k * m,
                    "d2h_bytes_per_step": n, "ms_per_step": e2e_s_max / e2e_steps * 1e3, "steps": e2e_steps,
                    "rank0_ms_of_each_step": e2e_step_ms,
                    "call": "alll_upload_fixedk(host) + alll_randomize + alll_solve + alll_get_assignment(host)",
                    "h2d_transport": {"host_buffer_bytes": 4 * k * m, "link_bytes_last_step": int(up["link_bytes"]),
                                      "chunks_packed": int(up["packed_chunks"]), "chunks_as_they_are": int(up["raw_chunks"]),
                                      "pack_threads": int(up["pack_threads"]),
                                      "note": "h2d_bytes_per_step counts the caller's host buffer (the tensor that is copied); the library's host "
                                              "threads re-pack its chunks to 25 bits per literal inside the timed region before they cross PCIe "
                                              "(include/alll_b200.h: alll_upload_info; ALLL_H2D_PACK=0 sends the words as they are)"}},
            "gpu_launches": int(launches_all),
            "clocks": clk,
            "between_sweeps_ms_per_solve": sum(s.between_sweeps_ms for s in stats) / args.steps,
            "incremental_mode": incremental,
            "other_workloads": extras,
            "sharded": sharded,
            "cfg5_multi_gpu": cfg5_multi,
        }
        emit_line(line)
    solver.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def other_workloads(device: int):
    """cfg2 (7-SAT n=1M m~4M) solve, cfg3 round-capped, cfg5 (8,192 x 5-SAT n=10k) batch, enumerated clauses and a ragged
    CSR instance on one GPU; each checked, each with a `roofline` object that names its bound honestly (only cfg4's main
    line is HBM-bound; the others are L2 / issue / latency bound and say so -- `peak` is always the measured HBM copy
    peak so that the fractions are comparable)."""
    import torch

    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_batch_torch, bounded_degree_ksat_torch, uniform_ksat_torch

    out = {}
    peak, _ = peaks()

    def roof(bound, achieved, note, **extra):
        return dict(bound=bound, achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak, note=note, **extra)

    c2 = CONFIGS["cfg2"]
    lits = bounded_degree_ksat_torch(c2["n"], c2["k"], c2["d"], INSTANCE_SEED_BASE + 2)
    m, k = int(lits.shape[0]), int(lits.shape[1])
    s = capi.Solver(device=device)
    s.upload_fixedk_device(c2["n"], m, k, lits.data_ptr())
    for i in range(3):
        s.randomize(50 + i)
        st = s.solve(50 + i)
    ok = st.status == 0 and independent_check(lits, s.get_assignment())
    sw_ms, _ = s.time_sweep(20)
    alg2 = 4 * k * m + c2["n"] // 8
    out["cfg2"] = {"workload": f"bounded-degree 7-SAT n={c2['n']} m={m}", "time_to_sat_ms": st.solve_ms, "sweeps": st.n_iterations,
                   "clause_evals_per_sec": st.n_clause_evals / (st.solve_ms * 1e-3), "verified": bool(ok), "sweep_ms": sw_ms,
                   "sweep_in_solve_ms": st.sweep_ms / max(st.n_iterations, 1), "between_sweeps_ms": st.between_sweeps_ms,
                   "luby_steps": st.n_luby_steps,
                   "sweep_algorithmic_GBps": alg2 / (sw_ms * 1e-3) / 1e9,
                   "roofline": roof("l2+issue", alg2 * st.n_iterations / (st.solve_ms * 1e-3) / 1e9,
                                    "whole solve launch (solve_persistent_kernel<7,7,7,5>): the 112 MB literal stream fits the 126 MB L2, so "
                                    "sweeps after the first read L2, not HBM -- the sweep body is bound by issue slots (ncu: profiles/r02_*cfg2*) "
                                    "and the solve by the independent-set phases between the sweeps",
                                    sweep_phase_GBps=alg2 / (st.sweep_ms / max(st.n_iterations, 1) * 1e-3) / 1e9 if st.sweep_ms else None,
                                    between_sweeps_share=st.between_sweeps_ms / st.solve_ms),
                   "note": "112 MB literal stream fits the 126 MB L2: back-to-back sweeps are L2-resident, not HBM-bound"}
    del lits
    c3 = CONFIGS["cfg3"]
    lits3 = uniform_ksat_torch(c3["n"], c3["k"], c3["m"], INSTANCE_SEED_BASE + 3)
    s.upload_fixedk_device(c3["n"], int(lits3.shape[0]), c3["k"], lits3.data_ptr())
    s.randomize(60)
    st3 = s.solve(60, 400)
    left, _ = s.eval(want_ids=False)
    alg3 = 4 * c3["k"] * c3["m"] + c3["n"] // 8
    out["cfg3"] = {"workload": f"uniform 3-SAT n={c3['n']} m={c3['m']} (ratio 3.0, beyond the LLL bound)", "round_cap": 400,
                   "us_per_round": st3.solve_ms * 1e3 / max(st3.n_iterations, 1),
                   "sweep_us_per_round": st3.sweep_ms * 1e3 / max(st3.n_iterations, 1),
                   "between_sweeps_us_per_round": st3.between_sweeps_ms * 1e3 / max(st3.n_iterations, 1),
                   "luby_steps_per_round": st3.n_luby_steps / max(st3.n_iterations, 1),
                   "roofline": roof("independent-set latency", alg3 * st3.n_iterations / (st3.solve_ms * 1e-3) / 1e9,
                                    "whole solve launch (solve_persistent_kernel<3,3,3,3>): 36 MB of literals and 16 MB of claim words are "
                                    "L2-resident; |U| stays at 1.4e5..3.8e5, so every round is a grid-wide Luby computation of 6-9 steps, each a "
                                    "grid barrier + two L2 round trips -- the sweep is ~10 % of a round (ncu: profiles/r02_*cfg3*)"),
                   "status": "MAX_ROUNDS" if st3.status == 1 else "OK", "time_to_sat_ms": None if st3.status == 1 else st3.solve_ms,
                   "rounds_per_sec": (st3.n_iterations - (0 if st3.status == 1 else 1)) / (st3.solve_ms * 1e-3),
                   "clause_evals_per_sec": st3.n_clause_evals / (st3.solve_ms * 1e-3), "violated_after_cap": left,
                   "note": "whole-clause Moser-Tardos resampling plateaus at this density (the reference does not terminate either)"}
    del lits3
    # SURVEY 8d's two optional companions: uniform 3-SAT at ratio 2.0 (the convergent counterpart of cfg3, same n) and
    # cfg4's strict-LLL variant (d = 12: k (d - 1) = 88 <= 2^8 / e - 1).  Never lose the other entries to these.
    try:
        m3b = 2_000_000
        lits3b = uniform_ksat_torch(c3["n"], c3["k"], m3b, INSTANCE_SEED_BASE + 13)
        s.upload_fixedk_device(c3["n"], m3b, c3["k"], lits3b.data_ptr())
        s.randomize(61)
        st3b = s.solve(61, 4000)
        ok3b = st3b.status == 0 and independent_check(lits3b, s.get_assignment())
        out["cfg3_convergent"] = {"workload": f"uniform 3-SAT n={c3['n']} m={m3b} (ratio 2.0)", "round_cap": 4000,
                                  "status": "OK" if st3b.status == 0 else "MAX_ROUNDS", "verified": bool(ok3b),
                                  "time_to_sat_ms": st3b.solve_ms if st3b.status == 0 else None, "sweeps": st3b.n_iterations,
                                  "us_per_round": st3b.solve_ms * 1e3 / max(st3b.n_iterations, 1),
                                  "rounds_per_sec": (st3b.n_iterations - (1 if st3b.status == 0 else 0)) / (st3b.solve_ms * 1e-3),
                                  "clause_evals_per_sec": st3b.n_clause_evals / (st3b.solve_ms * 1e-3)}
        del lits3b
    except Exception as e:
        out["cfg3_convergent"] = {"error": repr(e)}
    try:
        c4s = CONFIGS["cfg4"]
        lits4s = bounded_degree_ksat_torch(c4s["n"], c4s["k"], 12, INSTANCE_SEED_BASE + 14)
        m4s = int(lits4s.shape[0])
        s.upload_fixedk_device(c4s["n"], m4s, c4s["k"], lits4s.data_ptr())
        for i in range(3):
            s.randomize(90 + i)
            st4s = s.solve(90 + i)
        ok4s = st4s.status == 0 and independent_check(lits4s, s.get_assignment())
        alg4s = 4 * c4s["k"] * m4s + c4s["n"] // 8
        out["cfg4_strict_lll"] = {"workload": f"bounded-degree 8-SAT n={c4s['n']} m={m4s} d=12 (strict symmetric LLL: k(d-1)=88 <= 2^8/e-1)",
                                  "time_to_sat_ms": st4s.solve_ms, "sweeps": st4s.n_iterations, "verified": bool(ok4s),
                                  "clause_evals_per_sec": st4s.n_clause_evals / (st4s.solve_ms * 1e-3),
                                  "sweep_in_solve_ms": st4s.sweep_ms / max(st4s.n_iterations, 1), "between_sweeps_ms": st4s.between_sweeps_ms,
                                  "roofline": roof("hbm", alg4s * st4s.n_iterations / (st4s.solve_ms * 1e-3) / 1e9,
                                                   "whole solve launch (solve_persistent_kernel): 0.48 GB of literals per sweep, 4x the L2",
                                                   sweep_phase_GBps=alg4s / (st4s.sweep_ms / max(st4s.n_iterations, 1) * 1e-3) / 1e9 if st4s.sweep_ms else None)}
        del lits4s
    except Exception as e:
        out["cfg4_strict_lll"] = {"error": repr(e)}
    c5 = CONFIGS["cfg5"]
    n_inst = 8192
    off, blits = bounded_degree_batch_torch(n_inst, c5["n"], c5["k"], c5["d"], INSTANCE_SEED_BASE + 5)
    host = blits.cpu().numpy().view(np.uint32)
    s.batch_upload(c5["n"], c5["k"], off.numpy().astype(np.uint64), host)
    best, solved = None, 0
    for r in range(3):
        stats_b, _, _, ms = s.batch_solve(np.arange(r * n_inst, (r + 1) * n_inst, dtype=np.uint64), want_assignments=False)
        best = ms if best is None else min(best, ms)
        solved = int((stats_b["status"] == 0).sum())
    sweeps5 = float(stats_b["n_iterations"].sum())
    lit_bytes5 = 4 * c5["k"] * host.shape[0]
    out["cfg5"] = {"workload": f"{n_inst} x bounded-degree 5-SAT n={c5['n']} (m~{host.shape[0] // n_inst} each), one CTA per instance",
                   "batch_ms": best, "instances_per_sec": n_inst / (best * 1e-3), "solved": solved,
                   "mean_sweeps_per_instance": float(stats_b["n_iterations"].mean()),
                   "roofline": roof("issue+l2", lit_bytes5 * sweeps5 / n_inst / (best * 1e-3) / 1e9,
                                    "batch_solve_small_kernel<5>: algorithmic bytes = literal bytes x sweeps of every instance; the kernel is bound "
                                    "by issue slots and L2 round trips (5 jobs per SM), and the resident jobs' literals fall out of L2 between "
                                    "sweeps (ncu r01: DRAM 1.8x the literal bytes per batch)", literal_bytes_per_batch=lit_bytes5)}
    _, _, winner, pms = s.batch_solve(np.arange(n_inst, dtype=np.uint64), portfolio=True, want_assignments=False)
    out["cfg5_portfolio"] = {"workload": f"instance 0 x {n_inst} seeds, first-SAT device flag", "first_sat_ms": pms, "winner_seed_index": winner}
    # enumerated clauses on cfg4's shape: uniform 8-SAT, n=10M, m=40M, nothing stored but the assignment
    c4 = CONFIGS["cfg4"]
    m4 = c4["n"] * c4["d"] // c4["k"]
    s.upload_builtin_generator(capi.GEN_UNIFORM, c4["n"], m4, c4["k"], INSTANCE_SEED_BASE + 4, 0, cap_records=max(4096, 4 * m4 >> c4["k"]))
    s.randomize(70)
    gen_ms, _ = s.time_sweep(10)
    s.randomize(70)
    stg = s.solve(70, 2000)
    out["enumerated_cfg4_shape"] = {"workload": f"uniform 8-SAT n={c4['n']} m={m4}, clauses generated in the sweep kernel (Philox), never stored",
                                    "sweep_ms": gen_ms, "clause_evals_per_sec_sweep": m4 / (gen_ms * 1e-3), "time_to_sat_ms": stg.solve_ms,
                                    "sweeps": stg.n_iterations, "verified": bool(stg.status == 0 and s.verify()),
                                    "note": "bound by integer issue + scattered L2 lookups, not HBM"}
    # ragged CSR instance through the warp-cooperative CSR sweep (north_star item 1 for general DIMACS widths):
    # widths 3..8, m = 40 M, n = 10 M, kept in CSR form (ALLL_FLAG_FORCE_CSR; by default such input is padded onto planes)
    try:
        g = torch.Generator(device="cuda")
        g.manual_seed(INSTANCE_SEED_BASE + 9)
        m_r, n_r = 40_000_000, 10_000_000
        widths = torch.randint(3, 9, (m_r,), generator=g, device="cuda", dtype=torch.int64)
        off_t = torch.zeros(m_r + 1, dtype=torch.int64, device="cuda")
        off_t[1:] = torch.cumsum(widths, 0)
        n_lit = int(off_t[-1])
        lit_t = (torch.randint(0, n_r, (n_lit,), generator=g, device="cuda", dtype=torch.int32) * 2 +
                 torch.randint(0, 2, (n_lit,), generator=g, device="cuda", dtype=torch.int32))
        off_np, lit_np = off_t.cpu().numpy().astype(np.uint64), lit_t.cpu().numpy().view(np.uint32)
        del widths, off_t, lit_t
        torch.cuda.empty_cache()
        sc = capi.Solver(device=device, flags=capi.FLAG_FORCE_CSR)
        sc.upload_csr(n_r, off_np, lit_np)
        info = sc.layout_info()
        sc.randomize(80)
        csr_ms, csr_viol = sc.time_sweep(10)
        sc.randomize(80)
        stc = sc.solve(80, 300)
        alg_csr = 4 * n_lit + 8 * (m_r + 1)
        ok_csr = None
        if stc.status == 0:
            from oracle.oracle import Oracle
            ok_csr = bool(Oracle().verify(off_np, lit_np, sc.get_assignment()))      # independent CPU evaluation of every clause
        out["csr_ragged_40M"] = {"workload": f"ragged CSR: widths 3..8 uniform, m={m_r}, n={n_r}, L={n_lit} literals, ALLL_FLAG_FORCE_CSR",
                                 "kernel": "sweep_csr_warp_kernel<false> (assignment 1.25 MB: lookups through L2, no bucketing in CSR form)",
                                 "sweep_ms": csr_ms, "violated_per_sweep": csr_viol, "clause_evals_per_sec_sweep": m_r / (csr_ms * 1e-3),
                                 "bytes_read_per_sweep": info["literal_bytes"],
                                 "time_to_sat_ms": stc.solve_ms if stc.status == 0 else None, "sweeps": stc.n_iterations,
                                 "status": "OK" if stc.status == 0 else "MAX_ROUNDS(300)", "verified_on_cpu": ok_csr,
                                 "roofline": roof("hbm+l2-lookups", alg_csr / (csr_ms * 1e-3) / 1e9,
                                                  "algorithmic bytes = 4 L + 8 (m + 1) (literals + the caller's 64-bit offsets; the device form reads 4 L + "
                                                  "L / 8 start bits + 4 bytes per 128-literal chunk instead); every looked-up literal costs an L2 sector "
                                                  "of the un-bucketed 1.25 MB assignment, which bounds this kernel before HBM does",
                                                  algorithmic_bytes=alg_csr)}
        sc.close()
        del off_np, lit_np
    except Exception as e:                           # never lose the other entries to this one
        out["csr_ragged_40M"] = {"error": repr(e)}
    s.close()
    torch.cuda.empty_cache()
    out["dropin_cpp"] = dropin_cpp_records()
    return out


def dropin_cpp_records(timeout_s: float = 170.0):
    """The reference-facing C++ entry point end to end: tools/dropin_bench (compiled by build()) builds heap Clause objects the
    way example/main.cpp:149-178 does and times SATInstance::solve(vector<ClauseArray*>*) (SATInstance.h:60-66) -- flatten,
    upload, solve on the GPU, var_arr->vars written -- on cfg2 and cfg4; every result is re-checked clause by clause on the
    host through the public Clause API.  A process of its own (C++ caller, own CUDA context), outside every timed region of
    this file; the object graphs (40 M clauses: ~5 GB of host heap) are built before its timer starts."""
    exe = os.path.join(ROOT, "tools", "dropin_bench")
    if not os.path.exists(exe):
        return {"error": "tools/dropin_bench is not built (__graft_entry__.build() compiles it)"}
    recs = {}
    t_start = time.perf_counter()
    for name in ("cfg2", "cfg4"):
        c = CONFIGS[name]
        left = timeout_s - (time.perf_counter() - t_start)
        if left < 20:
            recs[name] = {"error": "skipped: time budget of the drop-in records used up"}
            continue
        try:
            pkg = os.path.join(ROOT, "alllsatisfiabilitysolver_b200")
            env = dict(os.environ, LD_LIBRARY_PATH=pkg + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""))
            p = subprocess.run([exe, "--n", str(c["n"]), "--k", str(c["k"]), "--d", str(c["d"]), "--steps", "3", "--gpus", "1"],
                               capture_output=True, text=True, timeout=left, env=env)
            line = [ln for ln in p.stdout.splitlines() if ln.startswith("{")]
            recs[name] = json.loads(line[-1]) if line else {"error": f"rc={p.returncode}: {p.stderr[-300:]}"}
        except Exception as e:
            recs[name] = {"error": repr(e)}
    recs["call"] = "SATInstance<uint32_t>::solve(vector<ClauseArray*>*) on heap Clause objects, one process, one GPU; solve_call_ms is the whole call"
    return recs


def cfg5_over_ranks(rank, world, local_rank, n_seeds=8192, n_inst=8192, reps=3):
    """Seed s / instance i on GPU s mod N.  Portfolio: one winner word in rank 0's memory, peer-mapped by every rank and
    claimed with a system-scope atomicCAS.  Batch: no exchange.  Device-timed, max over ranks."""
    import torch
    import torch.distributed as dist

    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_batch_torch, bounded_degree_ksat
    from alllsatisfiabilitysolver_b200.sharded import CudaPortfolioBackend, MultiGpuPortfolio, partition_round_robin

    c5 = CONFIGS["cfg5"]
    lits = bounded_degree_ksat(c5["n"], c5["k"], c5["d"], seed=INSTANCE_SEED_BASE + 5)
    pf = MultiGpuPortfolio(CudaPortfolioBackend(local_rank), rank, world)
    pf.upload(c5["n"], lits)
    runs = []
    for r in range(reps):
        res = pf.solve(np.arange(r * n_seeds, (r + 1) * n_seeds, dtype=np.uint64))
        ok = res["n_finished"] == 1 and res["assignment"] is not None
        if ok:
            v = res["assignment"].astype(bool)
            ok = bool((v[lits >> 1] ^ (lits & 1).astype(bool)).any(axis=1).all())
        runs.append((res["ms"], res["winner_rank"], ok))
    mine = partition_round_robin(n_inst, world, rank)
    off, blits = bounded_degree_batch_torch(len(mine), c5["n"], c5["k"], c5["d"], INSTANCE_SEED_BASE + 5 + 1000 * rank)
    s = capi.Solver(device=local_rank)
    s.batch_upload(c5["n"], c5["k"], off.numpy().astype(np.uint64), blits.cpu().numpy().view(np.uint32))
    best, solved = None, 0
    for r in range(reps):
        dist.barrier()
        st, _, _, ms = s.batch_solve(mine.astype(np.uint64) + np.uint64(r * n_inst), want_assignments=False)
        tmax = torch.tensor([ms], dtype=torch.float64, device="cuda")
        tsum = torch.tensor([float((st["status"] == 0).sum())], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        best = float(tmax) if best is None else min(best, float(tmax))
        solved = int(tsum)
    s.close()
    pf.be.solver.close()
    return {"portfolio": {"seeds": n_seeds, "first_sat_ms": min(x[0] for x in runs), "winner_ranks": [x[1] for x in runs],
                          "all_verified": all(x[2] for x in runs),
                          "flag": "one word in rank 0's HBM, CUDA-IPC mapped by all ranks, system-scope atomicCAS"},
            "batch": {"instances": n_inst, "batch_ms": best, "instances_per_sec": n_inst / (best * 1e-3), "solved": solved,
                      "parallelism": "instance i on GPU i mod N, no exchange"}}


def sharded_solves(args, shape, rank, world, local_rank, host_group):
    """Strong scaling: the workload instance split into contiguous clause ranges over all GPUs, fused NVLink exchange
    (records stored into every peer by the sweep, arrival flags, no NCCL call in the loop).  Measured three ways:
      (a) one process per GPU (CUDA IPC mappings; torch.distributed only moves the handles and is the barrier), full sweeps
          every round and with incremental re-evaluation; device-timed, max over ranks;
      (b) the same END TO END from host buffers: every rank uploads only ITS 1/N of the literals from pinned host memory
          (N PCIe links side by side), solves, rank 0 reads the assignment back; wall clock, max over ranks;
      (c) ONE process, ONE call (alll_multi_*: what SATInstance(..., n_threads) / the CLI's --gpus reach): rank 0 drives all
          N GPUs over peer access while the other ranks wait on a host (gloo) barrier with idle GPUs; device-timed and
          end to end from ONE host buffer;
    plus the round-1 NCCL all-gather variant for the record.  Every result is checked by the independent checker."""
    import torch
    import torch.distributed as dist

    from alllsatisfiabilitysolver_b200 import capi
    from alllsatisfiabilitysolver_b200.instances import bounded_degree_ksat_torch, uniform_ksat_torch
    from alllsatisfiabilitysolver_b200.sharded import CudaShardBackend, P2PShardedSolver, ShardedSolver, partition

    k, n = shape["k"], shape["n"]
    seed_inst = INSTANCE_SEED_BASE + int(args.workload[3:])            # the SAME instance on every rank
    lits = (bounded_degree_ksat_torch(n, k, shape["d"], seed_inst) if shape["kind"] == "bounded"
            else uniform_ksat_torch(n, k, shape["m"], seed_inst))
    m = int(lits.shape[0])
    lo, hi = partition(m, world)[rank]
    local = lits[lo:hi].contiguous()
    local_host_t = torch.empty(local.shape, dtype=local.dtype, pin_memory=True)
    local_host_t.copy_(local)
    full_host_t = None
    if rank == 0:                                                       # (c) needs the whole instance in ONE host buffer
        full_host_t = torch.empty(lits.shape, dtype=lits.dtype, pin_memory=True)
        full_host_t.copy_(lits)
    torch.cuda.synchronize()
    max_rounds = min(args.max_rounds, 1 << 19)

    def hbarrier():
        torch.cuda.synchronize()
        dist.barrier(group=host_group)

    def run(solver, randomize, reset):
        res = []
        for i in range(args.warmup + args.steps):
            reset()
            randomize(3000 + i)
            st = solver.solve(3000 + i, max_rounds)
            if i >= args.warmup:
                res.append(st)
        return res

    def summarize(res):
        return {"time_to_sat_ms": float(np.mean([r.solve_ms for r in res])),
                "sweeps_per_solve": float(np.mean([r.n_iterations for r in res])),
                "clause_evals_per_sec": float(np.sum([r.n_clause_evals for r in res]) / (np.sum([r.solve_ms for r in res]) * 1e-3)),
                "all_sat": all(r.status == 0 for r in res)}

    out = {"scaling": "strong", "m_clauses_total": m,
           "parallelism": f"{world} contiguous clause ranges, replicated bit-packed assignment"}

    # ---- (a) one process per GPU, persistent kernel per rank ----
    for key, flags in (("full_sweeps", 0), ("incremental", capi.FLAG_INCREMENTAL)):
        p2p = P2PShardedSolver(local_rank, rank, world, persistent=True, flags=flags)
        p2p.upload_range(n, local, m, lo)
        res = run(p2p, p2p.randomize, lambda: None)
        mine_np = p2p.get_assignment()
        mine = torch.from_numpy(mine_np).cuda()
        ref = mine.clone()
        dist.broadcast(ref, 0)
        checked = independent_check(lits, mine_np) if rank == 0 else True      # the full instance against rank 0's replica
        flags_t = torch.tensor([int(bool((ref == mine).all())), int(p2p.solver.verify()), int(checked)], device="cuda")
        dist.all_reduce(flags_t, op=dist.ReduceOp.MIN)
        rec = summarize(res)
        rec.update({"replicas_bit_identical": bool(flags_t[0].item()), "all_ranges_verified_on_gpu": bool(flags_t[1].item()),
                    "verified_by_independent_checker": bool(flags_t[2].item())})
        if flags:
            rec["incremental_rounds_per_solve"] = float(np.mean([r.n_incremental_rounds for r in res])) if hasattr(res[0], "n_incremental_rounds") else None
        out[key] = rec
        if key == "full_sweeps":
            # ---- (b) end to end from host buffers, every rank uploads only its own range ----
            e2e_out_t = torch.empty(n, dtype=torch.uint8, pin_memory=True)
            host_np = local_host_t.numpy().view(np.uint32)

            def e2e_step(i):
                p2p.upload_range(n, host_np, m, lo)
                p2p.randomize(6000 + i)
                st = p2p.solve(6000 + i, max_rounds)
                if rank == 0:
                    p2p.solver.get_assignment(e2e_out_t.numpy())
                return st

            e2e_step(-1)
            steps = max(1, min(args.steps, args.e2e_steps))
            hbarrier()
            t0 = time.perf_counter()
            ev = 0
            for i in range(steps):
                ev += e2e_step(i).n_clause_evals
            hbarrier()
            dt = time.perf_counter() - t0
            tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            out["e2e_sharded"] = {"value": ev / float(tmax), "unit": UNIT, "ms_per_step": float(tmax) / steps * 1e3, "steps": steps,
                                  "h2d_bytes_per_step_per_gpu": 4 * k * (hi - lo), "d2h_bytes_per_step": n,
                                  "call": "per rank: alll_upload_fixedk(host, own 1/N range) + alll_p2p_create/connect + alll_randomize + "
                                          "alll_solve_p2p; rank 0: alll_get_assignment(host)"}
        p2p.solver.close()
    out.update({"exchange": "fused into ONE persistent kernel per GPU: the sweep stores violated records into every peer over "
                            "NVLink (one coalesced block per warp flush per peer), count + arrival flag after the grid barrier, "
                            "every GPU waits for all flags and runs the identical independent set; no NCCL call, kernel boundary or "
                            "host round trip per round",
                "time_to_sat_ms": out["full_sweeps"]["time_to_sat_ms"], "sweeps_per_solve": out["full_sweeps"]["sweeps_per_solve"],
                "clause_evals_per_sec": out["full_sweeps"]["clause_evals_per_sec"],
                "replicas_bit_identical": out["full_sweeps"]["replicas_bit_identical"] and out["incremental"]["replicas_bit_identical"],
                "all_verified_by_independent_checker": out["full_sweeps"]["verified_by_independent_checker"] and out["incremental"]["verified_by_independent_checker"],
                "all_sat": out["full_sweeps"]["all_sat"] and out["incremental"]["all_sat"]})

    # ---- round-1 variant for the record: host-driven, NCCL all-gather of the records per round ----
    try:
        be = CudaShardBackend(local_rank)
        ss = ShardedSolver(be, rank, world)
        ss.upload_range(n, local, m, lo)
        res = run(ss, be.randomize, be.solver.reset_stats)
        be.solver.close()
        out["nccl_allgather_variant_time_to_sat_ms"] = float(np.mean([r.solve_ms for r in res]))
    except Exception as e:
        out["nccl_allgather_variant_time_to_sat_ms"] = repr(e)
    del local, lits
    torch.cuda.empty_cache()

    # ---- (c) ONE process, ONE call: rank 0 drives every GPU; the other ranks wait on the host with idle GPUs ----
    hbarrier()
    if rank == 0:
        try:
            one = {}
            full_np = full_host_t.numpy().view(np.uint32)
            out_np = torch.empty(n, dtype=torch.uint8, pin_memory=True).numpy()
            for key, flags in (("full_sweeps", 0), ("incremental", capi.FLAG_INCREMENTAL)):
                with capi.MultiSolver(list(range(world)), flags=flags) as ms:
                    ms.upload_fixedk(n, full_np)
                    info = ms.info()
                    res = []
                    for i in range(args.warmup + args.steps):
                        ms.randomize(3000 + i)
                        st = ms.solve(3000 + i, max_rounds)
                        if i >= args.warmup:
                            res.append(st)
                    rec = summarize(res)
                    rec["verified_by_independent_checker"] = independent_check(full_np.reshape(m, k), ms.get_assignment())
                    if flags:
                        rec["incremental_rounds_per_solve"] = float(np.mean([r.n_incremental_rounds for r in res]))
                    one[key] = rec
                    if key == "full_sweeps":
                        one["layout"] = info

                        def step(i):
                            ms.upload_fixedk(n, full_np)
                            ms.randomize(7000 + i)
                            st = ms.solve(7000 + i, max_rounds)
                            ms.get_assignment(out_np)
                            return st

                        step(-1)
                        steps = max(1, min(args.steps, args.e2e_steps))
                        t0 = time.perf_counter()
                        ev = sum(step(i).n_clause_evals for i in range(steps))
                        dt = time.perf_counter() - t0
                        one["e2e"] = {"value": ev / dt, "unit": UNIT, "ms_per_step": dt / steps * 1e3, "steps": steps,
                                      "h2d_bytes_per_step": 4 * k * m, "d2h_bytes_per_step": n,
                                      "call": "alll_multi_upload_fixedk(ONE host buffer; every GPU copies its own 1/N) + alll_multi_randomize + "
                                              "alll_multi_solve + alll_multi_get_assignment(host)"}
            one["how"] = ("alll_multi_* from rank 0's process alone: peer access instead of CUDA IPC, one host thread starts every GPU's "
                          "persistent solve kernel; the other ranks' processes wait on a gloo barrier")
            out["single_process"] = one
        except Exception as e:
            out["single_process"] = {"error": repr(e)}
    hbarrier()
    return out


_JSON_FD = None


def claim_stdout():
    """stdout carries exactly ONE line, the JSON result: fd 1 is pointed at stderr for everything else (NCCL prints its
    version banner to stdout at NCCL_DEBUG=VERSION and WARN; libraries and child processes inherit fd 1), and emit_line
    writes to the saved descriptor."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(CONFIGS))
    ap.add_argument("--scale", type=float, default=1.0, help="scale n (and m) of the workload; 1.0 = BASELINE size")
    ap.add_argument("--max-rounds", type=int, default=100000)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the brief cfg2 / cfg5 runs at N=1")
    ap.add_argument("--no-sharded", action="store_true", help="skip the clause-range sharded solves at N>1")
    ap.add_argument("--full-ref-timebox", type=float, default=120.0,
                    help="--impl reference: seconds allowed for ONE reference solve of the full workload instance (0 = skip)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print(f"note: warmup {args.warmup} < 3; timing rules ask for >= 3", file=sys.stderr)
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
