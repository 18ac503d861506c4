#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_digital_ads.py tests/test_gpu_kats.py tests/test_gpu_simple_market.py -m gpu -q -x 2>&1 | tail -5
python tools/bench_wide.py > $out/bench_wide_call19.json 2> $out/bench_wide_call19.err; cat $out/bench_wide_call19.json; tail -3 $out/bench_wide_call19.err
