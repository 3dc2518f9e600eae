"""TEST INFRASTRUCTURE — CPU oracle for the CARCA hot path (see carca_oracle.py).

Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of `bench.py` may import
this package.  The product package `carca_replication_b200` never does.
"""
