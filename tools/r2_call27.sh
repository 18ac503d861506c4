#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call27.log
tail -4 $out/pytest_call27.log
python tools/bench_wide.py > $out/bench_wide_call27.json 2> $out/bench_wide_call27.err; cat $out/bench_wide_call27.json
python tools/bench_configs.py --only C2-thread --steps 30 2>&1 | tail -3
python tools/bench_configs.py --only C2-queue --steps 30 2>&1 | tail -2
