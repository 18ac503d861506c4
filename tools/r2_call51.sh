#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_kats.py -m gpu -q -x 2>&1 | tail -4
timeout 600 python tools/fuzz_campaign3.py --first 9000 --count 200 --chains 0 2>&1 | tail -1 | cut -c1-200
