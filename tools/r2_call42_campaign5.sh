#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 2700 python tools/fuzz_campaign5.py --count 4000 > $out/fuzz_campaign5.log 2>&1
grep -c MISMATCH $out/fuzz_campaign5.log
grep MISMATCH $out/fuzz_campaign5.log | head -8 | cut -c1-400
tail -1 $out/fuzz_campaign5.log | cut -c1-700
