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
    assert lib.alll_abi_version() == 6


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
