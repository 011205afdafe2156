"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/alll_b200.h declares; the product never routes through the oracle; no silent CPU fallback."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "alll_b200.h")
PKG = os.path.join(ROOT, "alllsatisfiabilitysolver_b200")


@pytest.fixture(scope="module")
def built():
    subprocess.run(["make", "-s", "-C", os.path.join(PKG, "csrc")], check=True)
    return os.path.join(PKG, "liballl_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"ALLL_API\s+[\w\s\*]*?\b(alll_\w+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    from alllsatisfiabilitysolver_b200 import capi

    assert declared_symbols() == sorted(capi.SYMBOLS)


def test_library_exports_every_declared_symbol(built):
    lib = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    out = subprocess.run(["nm", "-D", "--defined-only", built], capture_output=True, text=True, check=True).stdout
    exported = sorted(re.findall(r" T (alll_\w+)", out))
    assert exported == declared_symbols()          # nothing else leaks out of the library
    assert lib.alll_abi_version() == 8


def test_host_packer_of_the_packed_transport(tmp_path):
    """csrc/hostpack.cpp (25 bits per literal for the H2D copy) is plain host code: its AVX2 path equals its scalar path byte for
    byte, the literals come back through the device kernel's arithmetic, nothing is written past the arrays."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "hostpack_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cpp", "hostpack_check.cpp"),
                    os.path.join(PKG, "csrc", "hostpack.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "hostpack ok" in out.stdout, out.stdout + out.stderr


def test_pack_threads_keep_the_ring_discipline_against_a_simulated_link(tmp_path):
    """csrc/packpipe.h (pack threads + 4-slot ring of the packed transport) against a simulated copy engine that reads a slot
    well after the chunk was issued: every literal arrives intact for any relative speed of packers and link, with and
    without the as-it-is fallback."""
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "packpipe_check")
    subprocess.run([cxx, "-O2", "-std=c++17", "-pthread", "-I", os.path.join(PKG, "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "cpp", "packpipe_check.cpp"), os.path.join(PKG, "csrc", "hostpack.cpp")], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "packpipe ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_library_is_sm100a_only(built):
    out = subprocess.run(["cuobjdump", "--list-elf", built], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_a_device(built):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from alllsatisfiabilitysolver_b200 import capi

    with pytest.raises(capi.AlllError) as e:
        capi.Solver()
    assert e.value.status == capi.CUDA_ERROR and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under the product package may import, link or open it."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                if re.search(r"(from|import)\s+oracle|alll_oracle|liballl_ref|oracle/", text):
                    # comments that merely name the checker are fine; code references are not
                    for line in text.splitlines():
                        if re.search(r"(from|import)\s+oracle|#include.*oracle|liballl_ref|liballl_oracle", line):
                            bad.append((f, line.strip()))
    assert not bad, bad


def test_headline_kernels_do_not_spill(built):
    """The sweep body runs at the register limit of a 512-thread CTA, and inside `solve_persistent_kernel` that limit covers
    the whole call tree: code added to an out-of-line independent-set function has pushed the sweep loop into local memory
    before (cfg4 solve 3.2 -> 5.7 ms, profiles/r02_l2_policy.md).  The built library must keep the headline variants free
    of spills: a stack frame beyond the 8 bytes the out-of-line calls need means the streaming loop lost registers."""
    import shutil

    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([exe, "-res-usage", built], capture_output=True, text=True, check=True).stdout
    usage = {m.group(1): (int(m.group(2)), int(m.group(3))) for m in re.finditer(r"Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)", out)}
    checked = 0
    for name, (reg, stack) in usage.items():
        # cfg4 (packed, two resident leading literals), cfg2 (7-SAT, assignment fully staged), cfg3 (3-SAT)
        headline = any(v in name for v in ("ILi8ELi2ELi3ELi5ELb1E", "ILi7ELi7ELi7ELi5ELb0E", "ILi3ELi3ELi3ELi3ELb0E"))
        if "solve_persistent_kernel" in name and headline:
            assert reg <= 128 and stack <= 8, f"{name}: REG {reg} STACK {stack}"
            checked += 1
        if "sweep_planes_kernel" in name and headline:
            assert reg <= 128 and stack == 0, f"{name}: REG {reg} STACK {stack}"
            checked += 1
    assert checked >= 6
