"""N-rank PCIe ceiling of the e2e (host-buffer) path: every rank moves one launch's outputs (118 MB
pinned D2H) and actions (26 MB pinned H2D) CONCURRENTLY with the other ranks, as bench.py's e2e
leg does at --gpus N; sustained over 20 back-to-back rounds between two barriers, max over ranks.
    torchrun --nproc-per-node N tools/pcie_peak_nrank.py   ->  one JSON line from rank 0
The e2e ceiling of N GPUs on this host = N * 6 553 600 env-steps / round time."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    d2h_bytes, h2d_bytes, rounds = 117964800, 26214400, 20
    dev_out = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    host_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    dev_in = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    host_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def round_trip():
        with torch.cuda.stream(s1):
            host_out.copy_(dev_out, non_blocking=True)
        with torch.cuda.stream(s2):
            dev_in.copy_(host_in, non_blocking=True)

    for _ in range(3):
        round_trip()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(rounds):
        round_trip()
    s1.synchronize(); s2.synchronize()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) * 1e-3 / rounds], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        sec = float(t.item())
        print(json.dumps({"n_gpus": world, "round_ms_max_over_ranks": sec * 1e3,
                          "d2h_GBps_per_gpu": d2h_bytes / sec / 1e9,
                          "d2h_GBps_aggregate": world * d2h_bytes / sec / 1e9,
                          "e2e_ceiling_env_steps_per_s": world * 6553600 / sec}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
