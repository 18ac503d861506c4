#!/bin/bash
# round 2, GPU call 6: engine1 output staging: tests + configs + ncu of the C4 static-schedule unit
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call6.log
tail -4 $out/pytest_call6.log
timeout 600 python tools/bench_configs.py --steps 20 --only thread > $out/bench_configs_call6.jsonl 2>&1
PHX_ENGINE1_STAGE=0 timeout 600 python tools/bench_configs.py --steps 20 --only C4-stackelberg-thread > $out/bench_configs_call6_nostage.jsonl 2>&1
python - <<'PY'
import json
for f in ("gpurun_out/bench_configs_call6.jsonl", "gpurun_out/bench_configs_call6_nostage.jsonl"):
    print(f)
    for ln in open(f):
        try: d = json.loads(ln)
        except Exception: print(ln.strip()[:200]); continue
        print(" ", d["config"], d["kernel"], round(d["ms_per_launch"], 4), "ms", f'{d["env_steps_per_s"]:.3e}', round(d["frac_of_measured_hbm_peak"], 3))
PY
ncu --set full --clock-control none --import-source on -f -k regex:phx_jit_step -s 3 -c 1 -o $out/prof_jit_c4_call6 \
    python tools/bench_configs.py --only C4-stackelberg-thread-jit --steps 3 > /dev/null 2>&1
ls -la $out/prof_jit_c4_call6.ncu-rep
