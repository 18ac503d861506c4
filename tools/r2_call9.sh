#!/bin/bash
# round 2, GPU call 9: tests (user program, rollout records); action-ring depth A/B
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call9.log
tail -6 $out/pytest_call9.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 1000 --warmup 50 --configs C2 --no-cpu-baseline > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err || tail -5 $out/bench_c2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c2_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"]*1e3, 3), "us  frac", round(d["roofline"]["frac"], 4), " e2e", f'{d["e2e"]["value"]:.3e}')
except Exception as e:
    print(tag, "unreadable", e)
PY
}
run ring16 PHX_SC_RING=16
run ring32 PHX_SC_RING=32
run ring16b PHX_SC_RING=16
run ring32b PHX_SC_RING=32
