#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
for f in test_gpu_user_program test_gpu_digital_ads test_gpu_stochastic_shuffle; do
  timeout 900 compute-sanitizer --tool racecheck --racecheck-report hazard --print-limit 6 --log-file $out/race_$f.log \
    python -m pytest tests/$f.py -m gpu -q --timeout 600 -k "not scale and not sampled" > $out/race_${f}_pytest.log 2>&1
  echo "== $f: $(tail -1 $out/race_${f}_pytest.log)"; grep -E "RACECHECK SUMMARY" $out/race_$f.log
  grep -m3 -E "hazard detected|in kernel|at .*\+0x" $out/race_$f.log | cut -c1-250
done
