#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 1000 --warmup 50 --configs C2 --no-cpu-baseline > $out/bench_c2_$tag.json 2> $out/bench_c2_$tag.err || tail -5 $out/bench_c2_$tag.err
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/bench_c2_{tag}.json").read().strip().splitlines()[-1])
    print(tag, "ms", round(d["ms_per_step"]*1e3, 3), "us  frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print(tag, "unreadable", e)
PY
}
run l0a0 X=1
run l1a0 PHX_LIB=$PWD/build/variants/libphx_l1a0.so
run l0a1 PHX_LIB=$PWD/build/variants/libphx_l0a1.so
run l1a1 PHX_LIB=$PWD/build/variants/libphx_l1a1.so
run l0a0b X=1
