#!/bin/bash
# compute-sanitizer over the block engine: memcheck + racecheck on the small golden tests
set -u
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck; do
  timeout 1200 compute-sanitizer --tool $tool --log-file $out/sanitizer_${tool}_wide.log \
    python -m pytest tests/test_gpu_digital_ads.py tests/test_gpu_simple_market.py tests/test_gpu_user_program.py -m gpu -q -x \
    -k "shipped_size or more_than_seven or (matches_python_handlers and 40)" > $out/sanitizer_${tool}_pytest.log 2>&1
  tail -2 $out/sanitizer_${tool}_pytest.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid" $out/sanitizer_${tool}_wide.log | sort | uniq -c | head -10
done
