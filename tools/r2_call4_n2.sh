#!/bin/bash
# round 2, GPU call 4 (2 GPUs): NCCL test of the packed gather + bench.py --gpus 2 (gather sub-record)
set -u
out=gpurun_out; mkdir -p $out
nvidia-smi -L > $out/n2_host.txt; nproc >> $out/n2_host.txt
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_reset_and_io.py -m gpu -x -q 2>&1 | tail -6 > $out/pytest_n2.log
cat $out/pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus 2 --steps 200 --warmup 20 > $out/bench_n2.json 2> $out/bench_n2.err
tail -c 600 $out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 C2", d["value"], d["ms_per_step"], d["roofline"]["frac"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("gather"), indent=1)[:2500])
for k, v in d["configs"].items():
    print(k, v.get("kernel"), v.get("envs_per_gpu"), v.get("ms_per_step"), (v.get("roofline") or {}).get("frac"), (v.get("e2e") or {}).get("value"), v.get("error"))
PY
