#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 2400 python tools/fuzz_campaign.py --first 1000 --count 2500 > $out/fuzz_campaign.log 2>&1
grep -c MISMATCH $out/fuzz_campaign.log
tail -6 $out/fuzz_campaign.log | cut -c1-600
