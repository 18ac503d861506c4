#!/bin/bash
# round 2, GPU call 15: dense kernel with hoisted invariants + lane-parallel merge
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_dense.py tests/test_gpu_kats.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --configs C2,C5 --no-cpu-baseline > $out/bench_call15.json 2> $out/bench_call15.err; tail -c 300 $out/bench_call15.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call15.json").read().strip().splitlines()[-1])
print("C2", d["ms_per_step"], round(d["roofline"]["frac"], 4))
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), v.get("error"))
PY
