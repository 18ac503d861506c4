#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_step -c 1 -o $out/prof_wide_call33 -f python tools/bench_wide.py --once --envs 4096 > $out/ncu_wide_call33.log 2>&1; tail -2 $out/ncu_wide_call33.log
