#!/bin/bash
# round 2, GPU call 5: tests (static schedule units); PDL A/B; engine configs with the static schedule
set -u
out=gpurun_out; mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call5.log
tail -4 $out/pytest_call5.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --configs C2 --no-cpu-baseline > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err || tail -5 $out/bench_c2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c2_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"]*1e3, 2), "us  frac", round(d["roofline"]["frac"], 3), " e2e", f'{d["e2e"]["value"]:.3e}', " single us", round(d["single_step"]["us_per_launch"], 2), d["config"]["timing"])
except Exception as e:
    print(tag, "unreadable", e)
PY
}
run pdl1 PHX_PDL=1
run pdl0 PHX_PDL=0
timeout 600 python tools/bench_configs.py --steps 20 --only thread > $out/bench_configs_call5.jsonl 2>&1
python - <<'PY'
import json
for ln in open("gpurun_out/bench_configs_call5.jsonl"):
    try: d = json.loads(ln)
    except Exception: print(ln.strip()[:200]); continue
    print(d["config"], d["kernel"], round(d["ms_per_launch"], 4), "ms", f'{d["env_steps_per_s"]:.3e}', round(d["frac_of_measured_hbm_peak"], 3))
PY
