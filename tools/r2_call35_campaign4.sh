#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 2700 python tools/fuzz_campaign4.py --count 120 > $out/fuzz_campaign4.log 2>&1
grep -c MISMATCH $out/fuzz_campaign4.log
grep MISMATCH $out/fuzz_campaign4.log | head -12 | cut -c1-500
tail -1 $out/fuzz_campaign4.log | cut -c1-900
