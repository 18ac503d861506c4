"""TEST INFRASTRUCTURE (CPU oracle).  Benchmark workloads written against the reference
plugin API.  Every builder takes the API module `ph` as its first argument, so the same
code runs on the unmodified reference (build container, through oracle/ref_shim.py) and
on oracle.phantom_oracle (anywhere)."""
