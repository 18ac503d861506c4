#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call39.log
tail -3 $out/pytest_call39.log
python tools/bench_wide.py > $out/bench_wide_call39.json 2> $out/bench_wide_call39.err; cat $out/bench_wide_call39.json
timeout 1200 compute-sanitizer --tool racecheck --log-file $out/sanitizer_racecheck_wide2.log \
    python -m pytest tests/test_gpu_digital_ads.py tests/test_gpu_simple_market.py tests/test_gpu_stochastic_shuffle.py tests/test_gpu_kats.py -m gpu -q -x \
    -k "shipped_size or more_than_seven or (stochastic_network_matches and wide) or wide_random" > $out/sanitizer_racecheck_wide2_pytest.log 2>&1
tail -2 $out/sanitizer_racecheck_wide2_pytest.log
grep -E "RACECHECK SUMMARY" $out/sanitizer_racecheck_wide2.log
timeout 2400 python tools/fuzz_campaign.py --first 8000 --count 400 > $out/fuzz_campaign_x.log 2>&1; tail -1 $out/fuzz_campaign_x.log | cut -c1-500
timeout 1200 python tools/fuzz_campaign3.py --first 3000 --count 600 --chains 0 > $out/fuzz_campaign3_x.log 2>&1; tail -1 $out/fuzz_campaign3_x.log | cut -c1-300
