#!/bin/bash
# final ncu captures of round 2: the C2, C4 and C5 kernels as shipped
set -u
out=gpurun_out; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sc_fast2 -c 1 -o $out/prof_final_c2 -f \
  python bench.py --configs C2 --steps 3 --warmup 1 --no-cpu-baseline > $out/ncu_final_c2.log 2>&1; tail -1 $out/ncu_final_c2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:phx_jit_step -c 1 -o $out/prof_final_c4 -f \
  python bench.py --configs C2,C4 --steps 3 --warmup 1 --sub-steps 3 --no-cpu-baseline > $out/ncu_final_c4.log 2>&1; tail -1 $out/ncu_final_c4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_step -c 1 -o $out/prof_final_c5 -f \
  python bench.py --configs C2,C5 --steps 3 --warmup 1 --sub-steps 3 --no-cpu-baseline > $out/ncu_final_c5.log 2>&1; tail -1 $out/ncu_final_c5.log
ls -la $out/prof_final_*.ncu-rep
