#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kats.py -m gpu -q -x 2>&1 | tail -4
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call36.log
tail -3 $out/pytest_call36.log
timeout 2400 python tools/fuzz_campaign.py --first 5000 --count 1500 > $out/fuzz_campaign_w.log 2>&1
grep -c MISMATCH $out/fuzz_campaign_w.log
grep MISMATCH $out/fuzz_campaign_w.log | head -5 | cut -c1-600
tail -1 $out/fuzz_campaign_w.log | cut -c1-800
