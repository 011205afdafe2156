"""The host side above the C ABI: drop-in C++ headers, the cnf_io loader and the CLI."""
import json
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "alllsatisfiabilitysolver_b200")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
LINK = ["-L" + PKG, "-lalll_b200", "-Wl,-rpath," + PKG, "-pthread"]


@pytest.fixture(scope="module")
def lib():
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "csrc")], check=True)
    return os.path.join(PKG, "liballl_b200.so")


@pytest.fixture(scope="module")
def cli(lib):
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "cli")], check=True)
    return os.path.join(PKG, "cli", "alll_solve")


def build(tmp, name, sources, extra=(), libs=()):
    out = os.path.join(tmp, name)
    subprocess.run([CXX, "-std=c++20", "-O1", "-w", *extra, *sources, "-o", out, *libs], check=True)
    return out


# ---- CPU -------------------------------------------------------------------------------------------

def test_reference_main_compiles_unchanged_against_dropin_headers(lib, tmp_path):
    """example/main.cpp of the reference, byte for byte, builds against include/ + cli/cnf_io of this repo
    (Boost replaced by the 100-line stub in tests/cpp/boost_stub, since Boost is not installed)."""
    ref_main = "/root/reference/example/main.cpp"
    if not os.path.exists(ref_main):
        pytest.skip("/root/reference is not present on this box")
    build(str(tmp_path), "ref_main", [ref_main, os.path.join(PKG, "cli", "cnf_io", "cnf_io.cpp")],
          extra=["-I" + os.path.join(ROOT, "tests", "cpp", "boost_stub"), "-I" + os.path.join(PKG, "include"),
                 "-I" + os.path.join(PKG, "cli")], libs=LINK)


def test_cnf_loader_matches_reference_parser_on_wellformed_files(tmp_path):
    """cnf_header_read / cnf_data_read return what the reference's cnf_io returned (tests/golden/dimacs/expected.json)."""
    exe = build(str(tmp_path), "cnf_dump", [os.path.join(ROOT, "tests", "cpp", "cnf_dump.cpp"),
                                            os.path.join(PKG, "cli", "cnf_io", "cnf_io.cpp")], extra=["-I" + os.path.join(PKG, "cli")])
    exp = json.load(open(os.path.join(GOLDEN, "dimacs", "expected.json")))
    for name in ("cfg1.cnf", "dialect.cnf"):
        got = json.loads(subprocess.run([exe, os.path.join(GOLDEN, "dimacs", name)], capture_output=True, text=True, check=True).stdout)
        assert got["error"] is False
        for key in ("v_num", "c_num", "l_num", "l_c_num", "l_val"):
            assert got[key] == exp[name][key], (name, key)


def test_cnf_loader_fixes_the_reference_hazards(tmp_path):
    """Deliberate deviations: last line without newline is parsed; tabs separate tokens;
    a SATLIB '%' trailer ends the data; contradictory counts and junk tokens are reported, never written out of bounds."""
    exe = build(str(tmp_path), "cnf_dump", [os.path.join(ROOT, "tests", "cpp", "cnf_dump.cpp"),
                                            os.path.join(PKG, "cli", "cnf_io", "cnf_io.cpp")], extra=["-I" + os.path.join(PKG, "cli")])

    def parse(text):
        p = tmp_path / "t.cnf"
        p.write_text(text)
        return json.loads(subprocess.run([exe, str(p)], capture_output=True, text=True, check=True).stdout)

    got = parse("p cnf 4 3\n1 2 0\n-3 4 0\n-1 -4 0")                      # reference drops the last clause
    assert got["error"] is False and got["l_c_num"] == [2, 2, 2] and got["l_val"] == [1, 2, -3, 4, -1, -4]
    got = parse("p cnf 3 1\n1\t-2\t3 0\n")                                 # reference reads only "1"
    assert got["error"] is False and got["l_val"] == [1, -2, 3]
    got = parse("c x\np cnf 2 2\n1 -2 0\n2 0\n%\n0\n")                      # reference writes l_c_num[c_num] out of bounds
    assert got["error"] is False and got["l_c_num"] == [2, 1]
    assert parse("p cnf 2 3\n1 -2 0\n2 0\n")["error"] is True              # header promises 3 clauses
    assert parse("p cnf 2 1\n1 x2 0\n")["error"] == "header"               # junk token
    assert parse("q cnf 2 1\n1 0\n")["error"] == "header"                  # no problem line
    assert parse("p cnf 2 1\n1 3 0\n")["error"] is True                    # variable beyond V


def test_cnf_loader_is_independent_of_how_the_file_is_cut_into_pieces(tmp_path):
    """The loader cuts the body at line starts into one piece per thread; clauses that span lines, comment lines,
    the '%' trailer and an unterminated tail must come out the same for every cut.  Also checks the cnf_read_csr
    extension (CSR + 2*var+neg encoding, example/main.cpp:168) against the cnf_io API result."""
    exe = build(str(tmp_path), "cnf_dump", [os.path.join(ROOT, "tests", "cpp", "cnf_dump.cpp"),
                                            os.path.join(PKG, "cli", "cnf_io", "cnf_io.cpp")], extra=["-I" + os.path.join(PKG, "cli")])
    rng = np.random.default_rng(5)
    lines, clauses, cur = ["c head", "p cnf 50 400"], [], []
    for _ in range(400):
        width = int(rng.integers(1, 9))
        cl = [int(v) * (1 if rng.random() < 0.5 else -1) for v in rng.integers(1, 51, size=width)]
        clauses.append(cl)
        toks = [str(x) for x in cl] + ["0"]
        while toks:                                           # break clauses over lines at random places
            take = int(rng.integers(1, len(toks) + 1))
            cur += toks[:take]
            toks = toks[take:]
            if rng.random() < 0.6:
                lines.append(("\t" if rng.random() < 0.2 else " ").join(cur))
                cur = []
                if rng.random() < 0.1:
                    lines.append("c 1 2 3 0 not a clause")
    lines.append(" ".join(cur + ["7", "-9"]))                  # unterminated tail: dropped
    lines += ["%", "0", "1 2 0"]                               # trailer: ignored
    path = tmp_path / "cut.cnf"
    path.write_text("\n".join(lines))                          # and no final newline
    exp_val = [x for cl in clauses for x in cl]
    exp_cnt = [len(cl) for cl in clauses]
    exp_lit = [2 * x - 2 if x > 0 else -2 * x - 1 for x in exp_val]
    exp_off = [0] + list(np.cumsum(exp_cnt))
    for threads, piece in ((1, 1 << 20), (3, 64), (8, 16), (64, 1)):
        env = dict(os.environ, ALLL_CNF_THREADS=str(threads), ALLL_CNF_PIECE_BYTES=str(piece))
        got = json.loads(subprocess.run([exe, str(path)], capture_output=True, text=True, check=True, env=env).stdout)
        assert got["error"] is False and got["l_c_num"] == exp_cnt and got["l_val"] == exp_val, (threads, piece)
        got = json.loads(subprocess.run([exe, str(path), "csr"], capture_output=True, text=True, check=True, env=env).stdout)
        assert got["error"] is False and got["off"] == exp_off and got["lit"] == exp_lit, (threads, piece)
        assert (got["v_num"], got["c_num"]) == (50, 400)


def test_user_program_and_cli_build(cli, lib, tmp_path):
    """Compile-and-link check of the reference-style user program (it runs in the GPU suite)."""
    build(str(tmp_path), "dropin_user", [os.path.join(ROOT, "tests", "cpp", "dropin_user.cpp")],
          extra=["-I" + os.path.join(PKG, "include")], libs=LINK)
    assert os.path.exists(cli)


def test_user_program_fails_loudly_without_a_device(lib, tmp_path):
    """No CPU fallback behind the drop-in headers either: without a CUDA device SATInstance::solve throws and says so."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    exe = build(str(tmp_path), "dropin_user", [os.path.join(ROOT, "tests", "cpp", "dropin_user.cpp")],
                extra=["-I" + os.path.join(PKG, "include")], libs=LINK)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode != 0 and "no CPU fallback" in out.stderr, out.stderr[-500:]


# ---- GPU -------------------------------------------------------------------------------------------

@pytest.mark.gpu
def test_cli_outputs_match_reference_format(cli, oracle, tmp_path):
    """Same observable behaviour as example/main.cpp: log lines, INFORMATION/STATISTICS blocks, six-field csv,
    'Variable i = b' dump, exit code 0 on a verified solution."""
    cnf = tmp_path / "cfg1.cnf"
    cnf.write_bytes(open(os.path.join(GOLDEN, "dimacs", "cfg1.cnf"), "rb").read())
    r = subprocess.run([cli, "-o", "-p", "4", "--sat", str(cnf), "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert re.search(r"^Log .*: Reading CNF file$", out, re.M)
    assert re.search(r"^Log .*: Read complete; Duration: [0-9.]+s$", out, re.M)
    assert "------------ INFORMATION ------------\n\t\t\t# Variables\t= 2000\n\t\t\t# Clauses\t= 1195\n" in out
    assert re.search(r"^Log .*: Starting parallel solve \(# Threads = 4\)$", out, re.M)
    assert re.search(r"^Log .*: Completed solve; Duration: [0-9.]+s$", out, re.M)
    m = re.search(r"# Iterations\t= (\d+)\n# Resamples\t= (\d+)\n\tThread 1: (\d+)\n\tThread 2: 0\n\tThread 3: 0\n\tThread 4: 0\n\nAvg. UNSAT MIS Size = (\d+)\n", out)
    assert m and m.group(2) == m.group(3)
    assert out.rstrip().endswith("SATISFIABLE")
    csv = (tmp_path / "cfg1.csv").read_text().strip().split(",")
    assert len(csv) == 6 and csv[1] == "2000" and csv[2] == "1195" and csv[4] == "4" and csv[5] == m.group(1)
    float(csv[0]); int(csv[3])
    dump = (tmp_path / "cfg1.out").read_text()
    assert dump.startswith(out)                                   # .out is a tee of stdout ...
    vals = re.findall(r"^Variable (\d+) = ([01])$", dump, re.M)     # ... plus the assignment
    assert [int(i) for i, _ in vals] == list(range(1, 2001))
    assign = np.array([int(b) for _, b in vals], np.uint8)
    exp = json.load(open(os.path.join(GOLDEN, "dimacs", "expected.json")))["cfg1.cnf"]
    assert oracle.check_signed(exp["l_c_num"], exp["l_val"], assign)   # independent checker (cnf_io.cpp:392-484 semantics)
    # same seed -> same run
    r2 = subprocess.run([cli, "--sat", str(cnf), "--seed", "5", "-p", "4"], capture_output=True, text=True)
    assert re.search(r"# Iterations\t= (\d+)", r2.stdout).group(1) == m.group(1)
    # unsatisfiable input + round cap -> exit code 1, no SATISFIABLE
    bad = tmp_path / "unsat.cnf"
    bad.write_text("p cnf 1 2\n1 0\n-1 0\n")
    r3 = subprocess.run([cli, "--sat", str(bad), "--max-rounds", "20"], capture_output=True, text=True)
    assert r3.returncode == 1 and "SATISFIABLE" not in r3.stdout.replace("UNSATISFIABLE", "")


@pytest.mark.gpu
def test_reference_style_user_program(lib, tmp_path):
    """A program written against the reference's public API only, built with the drop-in headers."""
    exe = build(str(tmp_path), "dropin_user", [os.path.join(ROOT, "tests", "cpp", "dropin_user.cpp")],
                extra=["-I" + os.path.join(PKG, "include")], libs=LINK)
    dimacs = tmp_path / "enum.cnf"
    r = subprocess.run([exe, str(dimacs)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    got = json.loads(r.stdout.strip().splitlines()[-1])
    assert got["valid"] and got["host_ok"] and got["host_ok2"] and got["status"] == 0
    assert got["thread_entries"] == 3 and got["thread0"] == got["resamples"]
    assert got["resamples"] % 5 == 0 and got["iterations"] >= 1 and got["iterations2"] >= 1
    head = dimacs.read_text().splitlines()
    assert head[0] == f"p cnf 3000 {got['n_clauses']}" and head[1].startswith(" ") and head[1].endswith(" 0")


@pytest.mark.gpu
def test_dropin_solve_over_several_gpus_from_cpp(lib, cli, oracle, tmp_path):
    """The reference's parallel-resource knob reaches more than one GPU from C++ (round-1 gap): SATInstance::set_gpus /
    ALLL_GPUS and the CLI's --gpus.  With a single GPU in the box the request is clamped to what is visible; the
    statistics and the assignment for a fixed seed do not depend on the number of GPUs."""
    import torch

    subprocess.run([os.path.join(ROOT, "tools", "build_dropin_bench.sh")], check=True, capture_output=True)
    exe = os.path.join(ROOT, "tools", "dropin_bench")
    visible = torch.cuda.device_count()
    runs = {}
    for gpus in sorted({1, min(2, visible), visible}):
        r = subprocess.run([exe, "--n", "200000", "--k", "7", "--d", "28", "--threads", "4", "--gpus", str(gpus), "--steps", "2"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        got = json.loads(r.stdout.strip().splitlines()[-1])
        assert got["all_ok"] and got["gpus_in_use"] == gpus, got
        runs[gpus] = got
    assert len({(g["m"], g["sweeps_per_solve"]) for g in runs.values()}) == 1     # same trajectory on 1..N GPUs
    # ALLL_GPUS for programs compiled unchanged (here: the reference-style user program), and the CLI flag
    exe2 = build(str(tmp_path), "dropin_user", [os.path.join(ROOT, "tests", "cpp", "dropin_user.cpp")],
                 extra=["-I" + os.path.join(PKG, "include")], libs=LINK)
    r = subprocess.run([exe2], capture_output=True, text=True, env=dict(os.environ, ALLL_GPUS="all"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert json.loads(r.stdout.strip().splitlines()[-1])["valid"]
    cnf = tmp_path / "cfg1.cnf"
    cnf.write_bytes(open(os.path.join(GOLDEN, "dimacs", "cfg1.cnf"), "rb").read())
    a = subprocess.run([cli, "--sat", str(cnf), "--seed", "5", "--gpus", "0"], capture_output=True, text=True)
    b = subprocess.run([cli, "--sat", str(cnf), "--seed", "5", "--gpus", "1"], capture_output=True, text=True)
    assert a.returncode == 0 and b.returncode == 0, a.stdout + a.stderr
    it = lambda t: re.search(r"# Iterations\t= (\d+)\n# Resamples\t= (\d+)", t).groups()
    assert it(a.stdout) == it(b.stdout) and a.stdout.rstrip().endswith("SATISFIABLE")


@pytest.mark.gpu
@pytest.mark.parametrize("ragged", [0, 1], ids=["uniform_streamed", "other_width_in_the_last_clause"])
def test_dropin_solve_streams_the_flatten_into_the_upload(lib, ragged):
    """SATInstance::solve on an instance large enough for the streamed path (>= 2^18 clauses): the flatten threads fill the
    page-locked staging buffer in order while alll_multi_upload_fixedk_streamed copies and lays out the chunks behind them.
    Same results as the non-streamed path; a clause of another width found at the very end of the flatten abandons the
    streamed upload and the general path takes over."""
    subprocess.run([os.path.join(ROOT, "tools", "build_dropin_bench.sh")], check=True, capture_output=True)
    exe = os.path.join(ROOT, "tools", "dropin_bench")
    base = [exe, "--n", "120000", "--k", "7", "--d", "28", "--threads", "4", "--steps", "2", "--ragged", str(ragged)]
    out = {}
    for mode, env in (("streamed", {}), ("whole_buffer", {"ALLL_NO_STREAMED_UPLOAD": "1"})):
        r = subprocess.run(base, capture_output=True, text=True, env={**os.environ, **env})
        assert r.returncode == 0, r.stdout + r.stderr
        out[mode] = json.loads(r.stdout.strip().splitlines()[-1])
        assert out[mode]["all_ok"] and out[mode]["m"] >= 1 << 18
    assert out["streamed"]["sweeps_per_solve"] == out["whole_buffer"]["sweeps_per_solve"]
