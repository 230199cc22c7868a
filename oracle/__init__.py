"""TEST INFRASTRUCTURE -- the CPU oracle of the DCFA-YOLO inference hot path.

Nothing under `dcfa-yolo_b200/` imports this package.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import, call or execute it, and only as the
checker.  See DESIGN.md section "Oracle" for how it is pinned to the reference.
"""
