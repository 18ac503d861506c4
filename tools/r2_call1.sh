#!/bin/bash
# round 2, GPU call 1: tests, bench (new kernels), A/B of the round-1 supply-chain kernel
set -u
out=gpurun_out; mkdir -p $out
nproc > $out/host.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> $out/host.txt; nvidia-smi -L >> $out/host.txt
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call1.log
tail -5 $out/pytest_call1.log
timeout 600 python bench.py --steps 200 --warmup 20 > $out/bench_call1.json 2> $out/bench_call1.err
tail -c 1500 $out/bench_call1.err
PHX_SC_V1=1 timeout 300 python bench.py --steps 200 --warmup 20 --configs C2 --no-cpu-baseline > $out/bench_call1_v1.json 2>> $out/bench_call1.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_call1.json","gpurun_out/bench_call1_v1.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "C2 ms", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "single", (d.get("single_step") or {}).get("us_per_launch"), d["config"]["timing"])
    for k,v in (d.get("configs") or {}).items():
        print("  ", k, v.get("kernel"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), (v.get("cpu_baseline") or {}).get("value"), v.get("error"))
    if "cpu_baseline" in d: print("  cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"], d["cpu_baseline"]["cores"])
PY
