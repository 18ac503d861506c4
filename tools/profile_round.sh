#!/bin/bash
# One GPU call that refreshes everything profiles/ cites for a round:
#   bash tools/profile_round.sh <tag>        (run under gpurun; outputs under gpurun_out/)
# 1. the driver's bench line, 2. ncu launch list of the same command, 3. one --set full capture
# of each dominant kernel (sc_fast_kernel / engine tile / engine thread / dense).
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
python bench.py > $out/bench_${tag}.json 2> $out/bench_${tag}.err
tail -c 3000 $out/bench_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file $out/launches_${tag}.csv python bench.py --steps 20 --warmup 3 --timed-only \
    > $out/bench_under_ncu_${tag}.log 2>&1
full="ncu --set full --clock-control none --import-source on -f"
$full -k regex:sc_fast -s 6 -c 1 -o $out/prof_sc_fast_${tag} \
    python bench.py --steps 8 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
$full -k regex:engine_step -s 3 -c 1 -o $out/prof_engine_c3_${tag} \
    python tools/bench_configs.py --only C3 --steps 3 > /dev/null 2>&1
$full -k regex:engine1_step -s 3 -c 1 -o $out/prof_engine1_c4_${tag} \
    python tools/bench_configs.py --only C4-stackelberg-thread --steps 3 > /dev/null 2>&1
$full -k regex:phx_jit_step -s 3 -c 1 -o $out/prof_jit_c4_${tag} \
    python tools/bench_configs.py --only C4-stackelberg-thread-jit --steps 3 > /dev/null 2>&1
$full -k regex:dense_step -s 3 -c 1 -o $out/prof_dense_c5_${tag} \
    python tools/bench_configs.py --only C5 --steps 3 > /dev/null 2>&1
python tools/bench_configs.py --steps 20 > $out/bench_configs_${tag}.jsonl 2>&1
ls -la $out | tail -12
