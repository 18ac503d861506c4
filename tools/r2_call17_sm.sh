#!/bin/bash
# round 2, GPU call 17: simple_market with up to 15 sellers (tile + block engines), env-level words on the block engine
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_simple_market.py tests/test_gpu_jit.py -m gpu -q 2>&1 | tail -40 > $out/pytest_call17_sm.log
tail -30 $out/pytest_call17_sm.log
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call17.log
tail -6 $out/pytest_call17.log
