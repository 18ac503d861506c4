#!/bin/bash
# round 2, GPU call 7: market collective resolve: tests + configs + ncu
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call7.log
tail -4 $out/pytest_call7.log
timeout 600 python tools/bench_configs.py --steps 10 --only C3 > $out/bench_configs_call7.jsonl 2>&1
timeout 600 python tools/bench_configs.py --steps 20 --only thread-jit >> $out/bench_configs_call7.jsonl 2>&1
PHX_COLLECTIVE=0 timeout 600 python tools/bench_configs.py --steps 10 --only C3-market >> $out/bench_configs_call7.jsonl 2>&1
python - <<'PY'
import json
for ln in open("gpurun_out/bench_configs_call7.jsonl"):
    try: d = json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print(" ", d["config"], d["kernel"], round(d["ms_per_launch"], 4), "ms", f'{d["env_steps_per_s"]:.3e}', round(d["frac_of_measured_hbm_peak"], 3))
PY
ncu --set full --clock-control none --import-source on -f -k regex:engine_step -s 3 -c 1 -o $out/prof_c3_call7 \
    python tools/bench_configs.py --only C3-market --steps 3 > /dev/null 2>&1
ls -la $out/prof_c3_call7.ncu-rep
