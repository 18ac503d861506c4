#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out
python tools/bench_wide.py > $out/bench_wide_call18.json 2> $out/bench_wide_call18.err; cat $out/bench_wide_call18.json; tail -3 $out/bench_wide_call18.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wide_step -c 1 -o $out/prof_wide_call18 -f python tools/bench_wide.py --once --envs 4096 > $out/ncu_wide_call18.log 2>&1; tail -3 $out/ncu_wide_call18.log
