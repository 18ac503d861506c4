#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call37.log
tail -3 $out/pytest_call37.log
timeout 600 python bench.py --configs C2,C3,C4,C5 --no-cpu-baseline > $out/bench_call37.json 2> $out/bench_call37.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call37.json").read().strip().splitlines()[-1])
print("C2", d["ms_per_step"], round(d["roofline"]["frac"], 4))
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), v.get("error"))
PY
