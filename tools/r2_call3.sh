#!/bin/bash
# round 2, GPU call 3: tests; hybrid e2e sweep; dense rewrite; ncu of sc_fast2 and dense
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > $out/pytest_call3.log
tail -4 $out/pytest_call3.log
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
run adaptive X=1
for k in 0 2 4 6 8 10 14; do run wire$k PHX_WIRE_CHUNKS=$k; done
run wire6_t8 PHX_WIRE_CHUNKS=6 PHX_HOST_THREADS=8
run wire6_t12 PHX_WIRE_CHUNKS=6 PHX_HOST_THREADS=12
timeout 300 python bench.py --steps 20 --warmup 5 --configs C2,C5 --sub-steps 20 --no-cpu-baseline > $out/bench_c5_call3.json 2> $out/bench_c5_call3.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_c5_call3.json").read().strip().splitlines()[-1])
c = d["configs"]["C5"]; print("C5", c["kernel"], c["ms_per_step"], c["roofline"]["frac"], c["e2e"]["value"])
PY
ncu --set full --clock-control none --import-source on -f -k regex:sc_fast2 -s 6 -c 1 -o $out/prof_sc_fast2_call3 \
    python bench.py --steps 8 --warmup 3 --timed-only --configs C2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -f -k regex:dense_step -s 2 -c 1 -o $out/prof_dense_call3 \
    python bench.py --steps 3 --warmup 3 --configs C5 --sub-steps 3 --no-cpu-baseline > /dev/null 2>&1
ls -la $out | tail -5
