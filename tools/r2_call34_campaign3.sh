#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 2700 python tools/fuzz_campaign3.py --count 2000 --chains 0 > $out/fuzz_campaign3.log 2>&1
grep -c MISMATCH $out/fuzz_campaign3.log
grep MISMATCH $out/fuzz_campaign3.log | head -10 | cut -c1-900
tail -1 $out/fuzz_campaign3.log | cut -c1-600
