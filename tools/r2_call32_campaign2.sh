#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 2400 python tools/fuzz_campaign2.py --market 40 --chain2 80 > $out/fuzz_campaign2_ads_market.log 2>&1
grep -c MISMATCH $out/fuzz_campaign2_ads_market.log
grep MISMATCH $out/fuzz_campaign2_ads_market.log | head -12 | cut -c1-330
tail -1 $out/fuzz_campaign2_ads_market.log | cut -c1-600
