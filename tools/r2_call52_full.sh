#!/bin/bash
# round 2, GPU call 13: full tests, smoke, both bench arms, launch list of the bench command
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call52.log
tail -4 $out/pytest_call52.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --impl reference > $out/bench_ref_call52.json 2> $out/bench_ref_call52.err
timeout 900 python bench.py > $out/bench_call52.json 2> $out/bench_call52.err; tail -c 400 $out/bench_call52.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_call52.csv \
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu_call52.log 2>&1
python - <<'PY'
import json
r = json.loads(open("gpurun_out/bench_ref_call52.json").read().strip().splitlines()[-1])
d = json.loads(open("gpurun_out/bench_call52.json").read().strip().splitlines()[-1])
print("REF", r["value"], r["cpu_baseline"]["kind"], "single", r["single_core"]["value"], {k: round(v.get("value", 0)) for k, v in r["configs"].items()})
print("OURS C2", f'{d["value"]:.4e}', "ms", d["ms_per_step"], "frac", round(d["roofline"]["frac"], 4), "e2e", f'{d["e2e"]["value"]:.4e}', "single", d["single_step"]["us_per_launch"], "full_io", d["full_io"]["ms_per_step"], round(d["full_io"]["frac"], 4), "launches", d["gpu_launches"])
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), "e2e", f'{(v.get("e2e") or {}).get("value", 0):.3e}', "cpu", round((v.get("cpu_baseline") or {}).get("value", 0)), v.get("error"))
PY
