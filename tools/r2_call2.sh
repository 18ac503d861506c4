#!/bin/bash
# round 2, GPU call 2: tests; A/B of the supply-chain kernels and of the compact wire e2e path; ncu
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call2.log
tail -4 $out/pytest_call2.log
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --configs C2 --no-cpu-baseline > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err || tail -5 $out/bench_c2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c2_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"]*1e3, 2), "us  frac", round(d["roofline"]["frac"], 3), " e2e", f'{d["e2e"]["value"]:.3e}', "threads", d["e2e"].get("host_threads"), " single us", round(d["single_step"]["us_per_launch"], 2))
except Exception as e:
    print(tag, "unreadable", e)
PY
}
run k3w4 PHX_SC_KERNEL=3 PHX_SC_WARPS=4
run k3w2 PHX_SC_KERNEL=3 PHX_SC_WARPS=2
run k2 PHX_SC_KERNEL=2
run k3w4_nowire PHX_SC_KERNEL=3 PHX_NO_WIRE=1
run k3w4_t8 PHX_SC_KERNEL=3 PHX_HOST_THREADS=8
run k3w4_t32 PHX_SC_KERNEL=3 PHX_HOST_THREADS=32
ncu --set full --clock-control none --import-source on -f -k regex:sc_fast3 -s 6 -c 1 -o $out/prof_sc_fast3_call2 \
    python bench.py --steps 8 --warmup 3 --timed-only --configs C2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:dense_step -s 2 -c 1 -o $out/prof_dense_call2 \
    python bench.py --steps 3 --warmup 3 --configs C5 --sub-steps 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $out | tail -8
