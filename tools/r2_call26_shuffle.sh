#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_stochastic_shuffle.py tests/test_gpu_supply_chain2.py -m gpu -q 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call26.log
tail -4 $out/pytest_call26.log
python tools/bench_wide.py > $out/bench_wide_call26.json 2> $out/bench_wide_call26.err; cat $out/bench_wide_call26.json
