"""TEST INFRASTRUCTURE ONLY -- the CPU oracle for the parallel Moser-Tardos path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``alllsatisfiabilitysolver_b200`` never does (tests/test_boundary.py checks that).
"""
