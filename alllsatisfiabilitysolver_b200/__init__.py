"""B200-native (sm_100a) parallel Moser-Tardos resampling path behind the reference's solver API.

* ``capi``      -- ctypes binding of the C ABI (``include/alll_b200.h`` -> ``liballl_b200.so``)
* ``instances`` -- seeded synthetic k-SAT generators and DIMACS helpers
* ``include/``  -- C++ drop-in headers mirroring ``library/include`` of the reference
* ``csrc/``     -- the CUDA kernels and the C ABI

No CPU fallback exists in this package; ``oracle/`` (the CPU checker) is test infrastructure.
"""
from . import capi, instances  # noqa: F401

__all__ = ["capi", "instances"]
