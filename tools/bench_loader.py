"""DIMACS loader throughput: this repository's cnf_io (alllsatisfiabilitysolver_b200/cli/cnf_io) against the reference's
(example/cnf_io/cnf_io.cpp, compiled where it lies; only possible where /root/reference exists).  CPU only.

    python tools/bench_loader.py [--config cfg2] [--scale 0.25]
"""
import argparse, json, os, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from alllsatisfiabilitysolver_b200 import instances  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--scale", type=float, default=0.25)
args = ap.parse_args()

work = os.path.join(ROOT, ".scratch", "loader")      # git- and gpurun-ignored
os.makedirs(work, exist_ok=True)
import numpy as np  # noqa: E402
cfg = instances.CONFIGS[args.config]
n, lits = instances.make_config(args.config, scale=args.scale)
path = os.path.join(work, f"{args.config}_{n}.cnf")
t0 = time.time()
sg = instances.lits_to_signed(lits)                      # vectorised writer (instances.write_dimacs is a per-clause loop)
sg = np.concatenate([sg, np.zeros((len(sg), 1), sg.dtype)], axis=1)
with open(path, "w") as f:
    f.write(f"c {args.config} scale {args.scale}\np cnf {n} {len(lits)}\n")
    np.savetxt(f, sg, fmt="%d")
size = os.path.getsize(path)
print(f"wrote {path}: {size / 1e6:.1f} MB, {len(lits)} clauses in {time.time() - t0:.1f} s", file=sys.stderr)

def build(out, inc, src):
    subprocess.check_call(["/usr/bin/g++", "-std=c++20", "-O2", "-o", out, os.path.join(ROOT, "tools", "loader_time.cpp"),
                           src, "-I", inc])

res = {"file_mb": size / 1e6, "clauses": int(len(lits)), "k": int(cfg["k"])}
ours = os.path.join(work, "loader_ours")
build(ours, os.path.join(ROOT, "alllsatisfiabilitysolver_b200", "cli"),
      os.path.join(ROOT, "alllsatisfiabilitysolver_b200", "cli", "cnf_io", "cnf_io.cpp"))
runs = [json.loads(subprocess.check_output([ours, path])) for _ in range(3)]
res["ours_ms"] = min(r["ms"] for r in runs)
res["ours_mb_per_s"] = size / 1e3 / res["ours_ms"]
ref_src = "/root/reference/example/cnf_io/cnf_io.cpp"
if os.path.exists(ref_src):
    ref = os.path.join(work, "loader_ref")
    build(ref, "/root/reference/example", ref_src)
    r = json.loads(subprocess.check_output([ref, path]))
    assert (r["v_num"], r["c_num"], r["l_num"], r["checksum"]) == tuple(runs[0][k] for k in ("v_num", "c_num", "l_num", "checksum"))
    res["reference_ms"] = r["ms"]
    res["reference_mb_per_s"] = size / 1e3 / r["ms"]
    res["speedup"] = r["ms"] / res["ours_ms"]
else:                                              # GPU box: the reference parser as compiled into oracle/_ref
    from oracle import oracle as orc
    if orc.have_reference():
        ref = orc.Reference()
        t0 = time.perf_counter()
        got = ref.cnf_read(path)
        res["reference_ms"] = (time.perf_counter() - t0) * 1e3
        assert got is not None and int(got[1]) == runs[0]["c_num"] and int(got[2]) == runs[0]["l_num"]
        res["reference_mb_per_s"] = size / 1e3 / res["reference_ms"]
        res["speedup"] = res["reference_ms"] / res["ours_ms"]
for t in (1, 2, 4, 8, 16, 32):
    if t > (os.cpu_count() or 1):
        break
    r = json.loads(subprocess.check_output([ours, path], env=dict(os.environ, ALLL_CNF_THREADS=str(t))))
    res.setdefault("ours_ms_by_threads", {})[t] = r["ms"]
res["host_cpus"] = os.cpu_count()
os.remove(path)
print(json.dumps(res))
