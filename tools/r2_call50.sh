#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python bench.py --no-cpu-baseline > $out/bench_call50.json 2> $out/bench_call50.err; tail -c 300 $out/bench_call50.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call50.json").read().strip().splitlines()[-1])
print("C2", d["ms_per_step"], round(d["roofline"]["frac"], 4), d["roofline"]["traffic"])
for k, v in d["configs"].items():
    print("  ", k, v.get("ms_per_step"), round(v["roofline"]["frac"], 4), v["roofline"].get("traffic"), v.get("error"))
PY
