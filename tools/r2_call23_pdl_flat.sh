#!/bin/bash
# (a) block engine with the flattened round queue: tests + timing; (b) sc_fast2 with a LATE programmatic-dependent-launch trigger
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_digital_ads.py tests/test_gpu_kats.py tests/test_gpu_simple_market.py tests/test_gpu_user_program.py -m gpu -q -x 2>&1 | tail -3
python tools/bench_wide.py > $out/bench_wide_call23.json 2> $out/bench_wide_call23.err; cat $out/bench_wide_call23.json; tail -3 $out/bench_wide_call23.err
for cfg in "0 0" "1 4" "1 8" "1 12" "1 20" "1 200"; do
  set -- $cfg
  PHX_PDL=$1 PHX_PDL_LEAD=$2 timeout 300 python bench.py --configs C2 --no-cpu-baseline > $out/bench_pdl_$1_$2.json 2> $out/bench_pdl_$1_$2.err
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(f"gpurun_out/bench_pdl_{sys.argv[1]}_{sys.argv[2]}.json").read().strip().splitlines()[-1])
    print("PDL", sys.argv[1], "lead", sys.argv[2], "ms", d["ms_per_step"], "frac", round(d["roofline"]["frac"], 4))
except Exception as e:
    print("PDL", sys.argv[1:], "failed", e)
PY
done
PHX_PDL=1 PHX_PDL_LEAD=8 timeout 600 python -m pytest tests/test_gpu_supply_chain.py -m gpu -q -x 2>&1 | tail -2
