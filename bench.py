#!/usr/bin/env python
"""bench.py -- env-steps/s of the supply-chain hot path (BASELINE config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: 65 536 parallel supply-chain envs per GPU (8 agent slots, 7 used), 100-step
episodes, synthetic uniform actions.  One bench "step" = ONE launch of the fused step
kernel through the C ABI (`phx_rollout`, T = 100 env transitions per env, auto-reset at the
episode boundary) = 6 553 600 env-steps per GPU.  See DESIGN.md "Measurement".

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline`,
`cpu_baseline`, `e2e`, `clocks`, `gpu_launches` are described in DESIGN.md.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

E_PER_GPU = 65536
T_EPISODE = 100
SEED = 0
# algorithmic HBM bytes: per env-step I/O and per-launch state traffic
B_IO = 4 + 12 + 4 + 2  # action f32, obs 3 x f32, reward f32, all_done 2 x u8
B_STATE = 2 * (8 + 16)  # header (step, episode) + shop state, read once + written once


# ------------------------------------------------------------------------------ helpers
def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the bench kernel, from the
    committed ncu --set full capture (profiles/traffic.json names the source file)."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t["source"]
    except Exception:
        return None, None


def port_calibration():
    """port speed / unmodified-reference speed, measured where both can run (the build
    container; tools/calibrate_port.py)."""
    try:
        with open(os.path.join(REPO, "tests", "golden", "port_calibration.json")) as f:
            return float(json.load(f)["port_over_reference"])
    except Exception:
        return None


WORKLOAD = ("supply-chain C2: 65536 envs/GPU x 8 agent slots (7 used), 100-step episodes; "
            "one bench step = one phx_rollout launch = 100 env transitions per env, "
            "auto-reset at the episode end")


def measured_peak_gbs():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            time.sleep(0.15)
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for name, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # the busy samples are the upper half of the distribution
        top = sorted(sm)[len(sm) // 2:]
        return {"sm_mhz": float(np.median(top)), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------- CPU baseline (port)
def _cpu_worker(args):
    """Steps one oracle env (the Python restatement of the reference loop, which runs at the
    reference's own speed: same object graph, same per-message work) for `seconds`."""
    worker, seconds = args
    import oracle.phantom_oracle as po
    from oracle import harness, rng
    from oracle.workloads import supply_chain as wl

    st = wl.order_stream(SEED, worker)
    env = wl.build(po, st)
    clock = harness.EpisodeClock([st])
    acts = np.random.RandomState(worker).uniform(0, 100, size=(T_EPISODE, 1)).astype(np.float32)
    clock.on_reset(); env.reset()
    for t in range(20):  # warm-up
        clock.on_step(env); env.step({"SHOP": acts[t]})
    clock.on_reset(); env.reset()
    n, t0 = 0, time.perf_counter()
    while True:
        for t in range(T_EPISODE):
            clock.on_step(env)
            env.step({"SHOP": acts[t]})
        n += T_EPISODE
        clock.on_reset(); env.reset()
        if time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def cpu_baseline(seconds: float, cores: int):
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(w, seconds) for w in range(cores)])
        wall = time.perf_counter() - t0
    steps = sum(n for n, _ in res)
    rate = sum(n / dt for n, dt in res)
    cal = port_calibration()
    return {
        "value": rate, "unit": "env-steps/s", "cores": cores, "kind": "port",
        "sample": (f"{cores} processes x {seconds:.0f} s of 100-step supply-chain episodes on "
                   f"oracle.phantom_oracle (Python restatement of PhantomEnv.step; the "
                   f"reference itself is Python and cannot travel to the GPU box), "
                   f"{steps} env-steps, wall {wall:.1f} s"),
        # the port runs this loop `port_over_reference` times as fast as the unmodified
        # reference on one core of the build container (tests/golden/port_calibration.json)
        "port_over_reference": cal,
    }


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from phantom_b200.envs.supply_chain import SupplyChainEnv

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # keep each rank (and the pinned host buffers it first-touches) on the CPUs next to its
        # GPU: with several ranks per box the host side of the e2e copies is NUMA-sensitive
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
            cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (m >> b) & 1}
            cpus &= set(os.sched_getaffinity(0))
            if cpus:
                os.sched_setaffinity(0, cpus)
        except Exception:
            pass
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E, T = E_PER_GPU, T_EPISODE
    env = SupplyChainEnv(num_envs=E, seed=SEED, device=local, env_offset=rank * E,
                         auto_reset=True)
    env.reset_batch()

    # inputs larger than L2: NBUF independent action / output sets (each 26 + 118 MB) are
    # cycled, so a launch never finds its inputs or output lines in the 126 MB L2
    NBUF = 4
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = [torch.rand((T, E, 1, 1), generator=gen, device=dev) * 100.0 for _ in range(NBUF)]
    outs = [env._alloc_outputs((T,)) for _ in range(NBUF)]
    # the base env never terminates agents early: the four per-agent mask planes are
    # constant, so the trainer-facing outputs are obs, reward and all_done
    from phantom_b200 import _lib as L
    import ctypes as C

    stream = torch.cuda.current_stream(dev).cuda_stream

    def launch(i):
        o = outs[i % NBUF]
        L.check(L.lib.phx_rollout(env._handle, T, acts[i % NBUF].data_ptr(), None,
                                  o.observations.data_ptr(), None, o.rewards.data_ptr(), None,
                                  None, None, o.all_done.data_ptr(), stream))

    for i in range(args.warmup):
        launch(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()

    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        ev0.record()
        for i in range(args.steps):
            launch(i)
        ev1.record()
        torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    env.check_errors()

    # ---- single-step API (SURVEY 8(d) asks for both): T = 1 phx_step calls, 100 of them
    # captured into one CUDA graph (a trainer that steps every env once per policy forward),
    # lean outputs like the rollout above; each launch moves 4.7 MB, so this leg is
    # launch/latency-bound and L2-resident by construction -- reported, not the headline
    single = None
    if rank == 0 and not args.timed_only:
        a1, o1 = acts[0], outs[0]
        side = torch.cuda.Stream(dev)

        def step_once(t, st):
            L.check(L.lib.phx_step(env._handle, a1[t].data_ptr(), None,
                                   o1.observations[t].data_ptr(), None, o1.rewards[t].data_ptr(),
                                   None, None, None, o1.all_done[t].data_ptr(), st))

        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for t in range(3):
                step_once(t, side.cuda_stream)
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            for t in range(T):
                step_once(t, torch.cuda.current_stream(dev).cuda_stream)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize()
        reps = 50
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(reps):
            graph.replay()
        g1.record()
        torch.cuda.synchronize()
        us = g0.elapsed_time(g1) * 1e3 / (reps * T)
        single = {"api": "phx_step, T=1, 100 launches per CUDA graph", "us_per_launch": us,
                  "value": E / (us * 1e-6), "unit": "env-steps/s",
                  "bytes_per_launch": E * (B_IO + B_STATE)}
        env.check_errors()
        env.reset_batch()

    # ---- e2e: same metric through the host-buffer C-ABI call (pinned host memory, H2D of
    # the actions and D2H of obs / reward / all_done inside the timed region)
    if args.timed_only:  # profiling aid: only the device-timed region (see tools/profile_round.sh)
        if rank == 0:
            print(json.dumps({"timed_only": True, "ms_per_step": ms / args.steps,
                              "gpu_launches": args.steps}), flush=True)
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return
    e2e_steps = max(3, min(args.steps, 12))
    h_act = torch.empty((T, E, 1, 1), dtype=torch.float32).pin_memory()
    h_act.copy_(acts[0].cpu())
    h_obs = torch.empty((T, E, 1, 3), dtype=torch.float32).pin_memory()
    h_rew = torch.empty((T, E, 1), dtype=torch.float32).pin_memory()
    h_all = torch.empty((T, E, 2), dtype=torch.uint8).pin_memory()

    def launch_host():
        L.check(L.lib.phx_rollout_host(env._handle, T, h_act.data_ptr(), None, h_obs.data_ptr(),
                                       None, h_rew.data_ptr(), None, None, None,
                                       h_all.data_ptr()))

    for _ in range(2):
        launch_host()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        launch_host()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    env.check_errors()
    h2d = h_act.numel() * 4
    d2h = h_obs.numel() * 4 + h_rew.numel() * 4 + h_all.numel()

    env_steps_per_launch = E * T
    value = world * env_steps_per_launch * args.steps / (ms * 1e-3)
    e2e_value = world * env_steps_per_launch * e2e_steps / e2e_s

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        traffic, traffic_src = ncu_traffic()
        bytes_per_launch = E * (T * B_IO + B_STATE)
        launch_ms = ms / args.steps
        achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
        line = {
            "metric": "env-steps/sec (65k parallel supply-chain envs)",
            "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": launch_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {
                "workload": WORKLOAD,
                "envs_per_gpu": E, "transitions_per_launch": T, "kernel": env.exec_name,
                "l2": f"inputs larger than L2: {NBUF} rotating action/output sets of "
                      f"{(acts[0].numel() * 4 + bytes_per_launch) / 1e6:.0f} MB",
                "parallelism": f"env-sharded x{world}, no data-path collective",
            },
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_per_launch,
                "kernel": "sc_fast_kernel<5,false>",
            },
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "phx_rollout_host (pinned host buffers)",
                    # the link that bounds it: device->host rate of one rank's result rows
                    # (profiles/r01_pcie_peak.json: plain pinned D2H of this size = 56.3 GB/s)
                    "bound": "pcie d2h",
                    "d2h_GBps_per_gpu": d2h * e2e_value / (world * E * T) / 1e9},
            "single_step": single,
            # kernels of ours inside the device-timed region (one sc_fast_kernel per bench step);
            # the e2e region launches one kernel per pipeline chunk (8 per call)
            "gpu_launches": args.steps,
            "gpu_launches_e2e": e2e_steps * 8,
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            line["cpu_baseline"] = cpu_baseline(args.cpu_seconds, cores)
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------- reference arm
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    # each "step" = a bounded sample: all cores stepping 100-step episodes for cpu_seconds
    t0 = time.perf_counter()
    rates = []
    for _ in range(max(1, args.warmup and 1)):
        cpu_baseline(min(2.0, args.cpu_seconds), cores)
    for _ in range(max(1, min(args.steps, 3))):
        rates.append(cpu_baseline(args.cpu_seconds, cores))
    best = max(rates, key=lambda r: r["value"])
    value = float(np.mean([r["value"] for r in rates]))
    line = {
        "impl": "reference", "metric": "env-steps/sec (65k parallel supply-chain envs)",
        "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": len(rates),
        "warmup": 1, "ms_per_step": args.cpu_seconds * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "cpu": f"same dynamics on the host: one env object per process, {cores} "
                          f"processes, 100-step episodes incl. reset; each step = a "
                          f"{args.cpu_seconds:.0f} s sample"},
        "cpu_baseline": dict(best, value=value),
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timed-only", action="store_true",
                    help="profiling aid: run only the warm-up and the device-timed launches")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
