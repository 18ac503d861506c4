#!/bin/bash
# round 2, GPU call 14 (8 GPUs): the driver's N=8 bench command + the 2-GPU NCCL test
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L | wc -l > $out/n8_host.txt; nproc >> $out/n8_host.txt
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
s=$(date +%s)
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
   bench.py --gpus 8 --steps 200 --warmup 20 > $out/bench_n8.json 2> $out/bench_n8.err
echo "N=8 wall $(( $(date +%s) - s )) s"
tail -c 800 $out/bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 C2", f'{d["value"]:.4e}', d["ms_per_step"], round(d["roofline"]["frac"], 4), "e2e", f'{d["e2e"]["value"]:.4e}', "threads", d["e2e"]["host_threads"])
for k, v in (d.get("gather") or {}).items():
    print("  gather", k, v if not isinstance(v, dict) else (round(v["us_per_launch"], 1), f'{v["value_with_gather"]:.3e}', round(v["rx_GBps"], 1)))
for k, v in d["configs"].items():
    print("  ", k, v.get("kernel"), v.get("envs_per_gpu"), v.get("ms_per_step"), round((v.get("roofline") or {}).get("frac", 0), 4), f'{v.get("value", 0):.3e}', "e2e", f'{(v.get("e2e") or {}).get("value", 0):.3e}', v.get("error"))
PY
