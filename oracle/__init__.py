"""CPU oracle for the FREUD SAE hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``freud_b200/`` may import this package.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs use it, and only as the checker / CPU baseline.

Parity pinning: the reference (ksadov/FREUD) ships no tests, golden vectors or
fixtures, so this restatement is pinned against outputs
of the reference's own Python imported in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``).
"""
