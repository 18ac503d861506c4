#!/bin/bash
# compute-sanitizer memcheck over the whole GPU suite (as far as it gets in 25 minutes), then
# racecheck over the kernels that use shared-memory staging (supply chain, dense, engine1 staged)
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 compute-sanitizer --tool memcheck --log-file $out/sanitizer_memcheck_all.log \
  python -m pytest tests -m gpu -q --timeout 600 > $out/sanitizer_memcheck_all_pytest.log 2>&1
tail -3 $out/sanitizer_memcheck_all_pytest.log
grep -E "ERROR SUMMARY|Invalid|misaligned" $out/sanitizer_memcheck_all.log | sort | uniq -c | head
timeout 1200 compute-sanitizer --tool racecheck --log-file $out/sanitizer_racecheck_sel.log \
  python -m pytest tests/test_gpu_dense.py tests/test_gpu_jit.py tests/test_gpu_supply_chain.py -m gpu -q --timeout 900 \
  -k "not full_size and not exhaustive and not scale" > $out/sanitizer_racecheck_sel_pytest.log 2>&1
tail -3 $out/sanitizer_racecheck_sel_pytest.log
grep -E "RACECHECK SUMMARY|hazard" $out/sanitizer_racecheck_sel.log | sort | uniq -c | head
