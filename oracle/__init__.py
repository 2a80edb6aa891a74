"""ORACLE — test infrastructure only.

CPU restatements (numpy / torch-CPU fp32 / plain C) of the reference algorithms on the UCOD-DPL hot path, each
function citing the reference file:line it follows.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
CPU-baseline / `--impl reference` legs may import this package; the product (`ucod_dpl_b200`) never does.
"""
