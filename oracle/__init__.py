"""TEST INFRASTRUCTURE.  The oracle is the checker, never the product.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package.  ``smcpp_b200`` never does (tests/test_layout.py enforces it).
"""
