#!/bin/bash
# round 2, GPU call 11: tests; sc_fast4 (two lanes per env) vs sc_fast2
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call11.log
tail -4 $out/pytest_call11.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 1000 --warmup 50 --configs C2 --no-cpu-baseline > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err || tail -5 $out/bench_c2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c2_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"]*1e3, 3), "us  frac", round(d["roofline"]["frac"], 4), " e2e", f'{d["e2e"]["value"]:.3e}', " single us", round(d["single_step"]["us_per_launch"], 2))
except Exception as e:
    print(tag, "unreadable", e)
PY
}
run k2 PHX_SC_KERNEL=2
run k4r64 PHX_SC_KERNEL=4
run k4r32 PHX_SC_KERNEL=4 PHX_SC_RING=32
run k2b PHX_SC_KERNEL=2
run k4r64b PHX_SC_KERNEL=4
PHX_SC_KERNEL=4 ncu --set full --clock-control none --import-source on -f -k regex:sc_fast4 -s 6 -c 1 -o $out/prof_sc_fast4_call11 \
    python bench.py --steps 8 --warmup 3 --timed-only --configs C2 > /dev/null 2>&1
ls -la $out/prof_sc_fast4_call11.ncu-rep
