#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_user_program.py -m gpu -q -x 2>&1 | tail -15
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call20.log
tail -5 $out/pytest_call20.log
