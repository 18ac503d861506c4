#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_supply_chain.py tests/test_gpu_reset_and_io.py tests/test_gpu_vector.py tests/test_gpu_metrics.py tests/test_gpu_rollout_records.py -m gpu -q -x 2>&1 | tail -3
for pdl in 0 1; do
  PHX_PDL=$pdl timeout 300 python bench.py --configs C2 --no-cpu-baseline > $out/bench_pdl1_$pdl.json 2> $out/bench_pdl1_$pdl.err
  python - "$pdl" <<'PY'
import json, sys
d = json.loads(open(f"gpurun_out/bench_pdl1_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print("PDL", sys.argv[1], "ms", d["ms_per_step"], "frac", round(d["roofline"]["frac"], 4), "single-step us", d["single_step"]["us_per_launch"], "full_io", d["full_io"]["ms_per_step"])
PY
done
