"""pytest configuration: registers the ``gpu`` marker and shared fixtures.

``-m "not gpu"`` : oracle vs golden vectors / reference shim, host logic, C-ABI symbol checks (CPU only).
``-m gpu``       : parity tests proper -- the CUDA path through the C-ABI vs the oracle (needs a B200).
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 via gpurun)")


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle, build

    build(ref=os.path.isdir("/root/reference"))
    return Oracle()


@pytest.fixture(scope="session")
def reference(oracle):
    from oracle.oracle import Reference, have_reference

    if not have_reference():
        pytest.skip("oracle/_ref/liballl_ref.so not built (no /root/reference here)")
    return Reference()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden_v1.npz"))


def golden_case(golden, name):
    n = int(golden[f"{name}/n_vars"][0])
    return n, golden[f"{name}/off"], golden[f"{name}/lit"], golden[f"{name}/assign"]
