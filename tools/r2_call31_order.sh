#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call31.log
tail -4 $out/pytest_call31.log
timeout 2400 python tools/fuzz_campaign.py --first 1000 --count 2500 > $out/fuzz_campaign2.log 2>&1
grep -c MISMATCH $out/fuzz_campaign2.log
tail -1 $out/fuzz_campaign2.log | cut -c1-700
timeout 600 python bench.py --configs C2,C3,C4,C5 --no-cpu-baseline > $out/bench_call31.json 2> $out/bench_call31.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call31.json").read().strip().splitlines()[-1])
print("C2", d["ms_per_step"], round(d["roofline"]["frac"], 4))
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), v.get("error"))
PY
