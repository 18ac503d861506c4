#!/bin/bash
# round 2, GPU call 8: full tests; the driver's two bench commands, timed
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call8.log
tail -4 $out/pytest_call8.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
s=$(date +%s)
timeout 900 python bench.py --impl reference > $out/bench_ref_call8.json 2> $out/bench_ref_call8.err
echo "reference arm wall $(( $(date +%s) - s )) s"; tail -c 300 $out/bench_ref_call8.err
s=$(date +%s)
timeout 900 python bench.py > $out/bench_call8.json 2> $out/bench_call8.err
echo "our arm wall $(( $(date +%s) - s )) s"; tail -c 600 $out/bench_call8.err
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bench_ref_call8.json").read().strip().splitlines()[-1])
print("REF", r["value"], r["cpu_baseline"]["kind"], r["cpu_baseline"]["cores"], "single", r["single_core"]["value"], {k: v.get("value") for k, v in r["configs"].items()}, "wall", r["wall_s"])
d = json.loads(open("gpurun_out/bench_call8.json").read().strip().splitlines()[-1])
print("OURS C2", d["value"], "ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "single", d["single_step"]["us_per_launch"], d["config"]["timing"], d["clocks"])
print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], "ratio e2e", d["e2e"]["value"] / r["value"])
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), "e2e", (v.get("e2e") or {}).get("value"), "cpu", (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
PY
