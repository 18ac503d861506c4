#!/bin/bash
# final compute-sanitizer pass of round 2: memcheck over the whole GPU suite, racecheck over the kernels with
# shared-memory staging / queues (supply chain, dense, jit, block engine incl. shuffle, waiting mail)
set -u
out=gpurun_out; mkdir -p $out
timeout 1500 compute-sanitizer --tool memcheck --log-file $out/sanitizer_memcheck_final.log \
  python -m pytest tests -m gpu -q --timeout 900 > $out/sanitizer_memcheck_final_pytest.log 2>&1
tail -2 $out/sanitizer_memcheck_final_pytest.log
grep -E "ERROR SUMMARY|Invalid|misaligned" $out/sanitizer_memcheck_final.log | sort | uniq -c | head
timeout 1500 compute-sanitizer --tool racecheck --log-file $out/sanitizer_racecheck_final.log \
  python -m pytest tests/test_gpu_dense.py tests/test_gpu_jit.py tests/test_gpu_supply_chain.py tests/test_gpu_digital_ads.py \
     tests/test_gpu_stochastic_shuffle.py tests/test_gpu_kats.py tests/test_gpu_user_program.py -m gpu -q --timeout 900 \
  -k "not full_size and not exhaustive and not scale and not sampled" > $out/sanitizer_racecheck_final_pytest.log 2>&1
tail -2 $out/sanitizer_racecheck_final_pytest.log
grep -E "RACECHECK SUMMARY|hazard" $out/sanitizer_racecheck_final.log | sort | uniq -c | head
