#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_digital_ads.py -m gpu -q -x 2>&1 | tail -6
