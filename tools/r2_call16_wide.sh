#!/bin/bash
# round 2, GPU call 16: StageRule chains + the 128-lane block engine (full GPU suite)
set -u
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_digital_ads.py tests/test_gpu_kats.py -m gpu -q 2>&1 | tail -40 > $out/pytest_call16_wide.log
tail -30 $out/pytest_call16_wide.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call16.log
tail -6 $out/pytest_call16.log
