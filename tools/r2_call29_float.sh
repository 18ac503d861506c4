#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_kats.py -m gpu -q -x 2>&1 | tail -6
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $out/pytest_call29.log
tail -4 $out/pytest_call29.log
