#!/bin/bash
set -u
out=gpurun_out; mkdir -p $out; : > $out/pcie_nrank.jsonl
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
     tools/pcie_peak_nrank.py 2>/dev/null | tail -1 >> $out/pcie_nrank.jsonl
done
cat $out/pcie_nrank.jsonl; nproc; numactl -H 2>/dev/null | head -3
