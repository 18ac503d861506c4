#!/bin/bash
set -u
timeout 900 python -m pytest tests/test_gpu_kats.py -m gpu -q -x -k "waits or without_resolve" 2>&1 | tail -12
