#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python tools/fuzz_campaign6.py --first 7000 --count 30 > $out/fuzz_campaign6.log 2>&1
grep -c MISMATCH $out/fuzz_campaign6.log
grep MISMATCH $out/fuzz_campaign6.log | head -6 | cut -c1-500
tail -1 $out/fuzz_campaign6.log | cut -c1-700
