"""TEST INFRASTRUCTURE.

CPU oracle for the env-step hot path: a restatement of the reference's algorithm
(`phantom_oracle/`), the counter-based RNG contract (`rng.py`), the benchmark workloads
written against the reference plugin API (`workloads/`) and vectorised numpy restatements
for full-size checks (`vectorised.py`).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import anything from this package.  The product (phantom_b200/) never does.
"""
