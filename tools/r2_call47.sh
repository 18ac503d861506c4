#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call47.log
tail -3 $out/pytest_call47.log
timeout 600 python bench.py --configs C2,C4 --no-cpu-baseline > $out/bench_call47.json 2> $out/bench_call47.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_call47.json").read().strip().splitlines()[-1])
print("C2", d["ms_per_step"], round(d["roofline"]["frac"], 4))
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), v.get("error"))
PY
timeout 1500 compute-sanitizer --tool racecheck --log-file $out/sanitizer_racecheck_final.log \
  python -m pytest tests/test_gpu_dense.py tests/test_gpu_jit.py tests/test_gpu_supply_chain.py tests/test_gpu_digital_ads.py \
     tests/test_gpu_stochastic_shuffle.py tests/test_gpu_kats.py tests/test_gpu_user_program.py tests/test_gpu_stackelberg.py tests/test_gpu_simple_market.py -m gpu -q --timeout 900 \
  -k "not full_size and not exhaustive and not scale and not sampled" > $out/sanitizer_racecheck_final_pytest.log 2>&1
tail -1 $out/sanitizer_racecheck_final_pytest.log
grep -E "RACECHECK SUMMARY" $out/sanitizer_racecheck_final.log
