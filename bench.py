#!/usr/bin/env python
"""bench.py -- env-steps/s of the env-step hot path on the BASELINE configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--configs C2,C3,C4,C5]

Headline (the JSON line's top-level keys) = BASELINE config C2: 65 536 parallel supply-chain
envs per GPU (8 agent slots, 7 used), 100-step episodes, synthetic uniform actions.  One bench
"step" = ONE launch of the fused step kernel through the C ABI (`phx_rollout`, T = 100 env
transitions per env, auto-reset at the episode boundary) = 6 553 600 env-steps per GPU.  The K
timed launches are captured in one CUDA graph (no launch gaps inside the timed region).

`configs` holds one sub-record per further BASELINE config -- C3 (3-stage FSM market, 32 768
envs x 32 agents), C4 (Stackelberg, 131 072 envs x 4 agents, sharded over the N GPUs), C5
(dense graph, 16 384 envs x 128 agents, sharded) -- each with its own kernel, roofline,
cpu_baseline and e2e.  See DESIGN.md "Measurement".

Prints ONE JSON line (rank 0).  Keys follow the driver contract; `roofline`, `cpu_baseline`,
`e2e`, `clocks`, `gpu_launches`, `configs`, `gather` are described in DESIGN.md.
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

SEED = 0
REF_COPY = os.path.join(REPO, "baseline", "_ref")  # unmodified reference sources (git-ignored)

# ------------------------------------------------------------------------------ configs
# Algorithmic HBM bytes: bytes per launch = E * (T * b_io + b_state);
#   b_io    per env-step: actions + every output plane the launch writes
#   b_state per env and launch: persisted state, read once + written once
CONFIGS = {
    "C2": dict(
        workload=("supply-chain C2: 65536 envs/GPU x 8 agent slots (7 used), 100-step episodes; "
                  "one bench step = one phx_rollout launch = 100 env transitions per env, "
                  "auto-reset at the episode end"),
        E=65536, scaling="weak", T=100, S=1, O=3, lean=True, act_scale=100.0, binary_from=None,
        b_io=4 + 12 + 4 + 2, b_state=2 * (8 + 16), nbuf=4),
    "C3": dict(
        workload=("FSM market C3: 32768 envs/GPU x 32 agents (7 makers + 24 takers strategic + "
                  "clearing), 3 stages, 99-step episodes; one bench step = one phx_rollout launch "
                  "= 99 transitions per env, auto-reset"),
        E=32768, scaling="weak", T=99, S=31, O=3, lean=False, act_scale=1.0, binary_from=7,
        b_io=31 * (4 + 12 + 4 + 4) + 2, b_state=2 * (16 + 8 + 32 * (8 * 4 + 4 + 12) + 8), nbuf=2,
        specialise=True),
    "C4": dict(
        workload=("Stackelberg C4: 131072 envs in total (sharded over the GPUs) x 4 agents, "
                  "100-step episodes; one bench step = one phx_rollout launch = 100 transitions "
                  "per env, auto-reset"),
        E=131072, scaling="strong", T=100, S=4, O=2, lean=False, act_scale=1.0, binary_from=None,
        b_io=4 * (4 + 8 + 4 + 4) + 2, b_state=2 * (16 + 8 + 8 * (16 + 4) + 4), nbuf=2,
        # env.specialise(): the step kernel is rebuilt at run time (nvcc -cubin, cached) with this
        # env class as a compile-time constant and its STATIC message schedule (phantom_b200/jit.py)
        specialise=True),
    "C5": dict(
        workload=("dense graph C5: 16384 envs in total (sharded over the GPUs) x 128 agents, "
                  "complete graph, BatchResolver(round_limit=2), 16 256 + 128 messages per step, "
                  "8-step episodes; one bench step = one phx_rollout launch = 8 transitions per "
                  "env, auto-reset"),
        E=16384, scaling="strong", T=8, S=128, O=3, lean=False, act_scale=1.0, binary_from=None,
        b_io=128 * (4 + 12 + 4 + 4) + 2, b_state=2 * (16 + 6 * 128 * 4), nbuf=2),
}


def make_env(name, **kw):
    if name == "C2":
        from phantom_b200.envs.supply_chain import SupplyChainEnv
        return SupplyChainEnv(**kw)
    if name == "C3":
        from phantom_b200.envs.market import MarketEnv
        return MarketEnv(**kw)
    if name == "C4":
        from phantom_b200.envs.stackelberg_game import StackelbergGameEnv
        return StackelbergGameEnv(**kw)
    if name == "C5":
        from phantom_b200.envs.dense import DenseEnv
        return DenseEnv(**kw)
    raise ValueError(name)


# ------------------------------------------------------------------------------ helpers
def ncu_traffic(config=None):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the bench kernel (or of the
    kernel of sub-config `config`), from the committed ncu --set full capture
    (profiles/traffic.json names the source file).  (None, None) if there is no capture."""
    try:
        with open(os.path.join(REPO, "profiles", "traffic.json")) as f:
            t = json.load(f)
        if config is not None:
            t = t["configs"][config]
        return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t["source"]
    except Exception:
        return None, None


def pcie_ceiling(n_gpus, e2e_value):
    """The measured N-rank host-copy ceiling of the e2e path (profiles/pcie_nrank.json: every rank
    moves a launch's 118 MB out / 26 MB in, all ranks at once) and the e2e figure as a fraction
    of it.  {} if no measurement for this N is committed."""
    try:
        with open(os.path.join(REPO, "profiles", "pcie_nrank.json")) as f:
            t = json.load(f)
        ceil = float(t["per_n"][str(int(n_gpus))]["e2e_ceiling_env_steps_per_s"])
        return {"ceiling": ceil, "frac_of_ceiling": float(e2e_value) / ceil,
                "ceiling_source": t["source"]}
    except Exception:
        return {}


def port_calibration():
    """port speed / unmodified-reference speed, measured where both can run (the build
    container; tools/calibrate_port.py)."""
    try:
        with open(os.path.join(REPO, "tests", "golden", "port_calibration.json")) as f:
            return float(json.load(f)["port_over_reference"])
    except Exception:
        return None


def measured_peak_gbs():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed regions run."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "20"],
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


# ----------------------------------------------------------------- CPU arms (host cores)
class _NativeStream:
    """The reference draws from process-global numpy generators; the TIMED baselines keep that
    ('the timed baseline runs without the patch')."""

    def randint(self, n):
        return int(np.random.randint(n))


def _reference_module():
    """The UNMODIFIED reference, imported from baseline/_ref (a copy of /root/reference's
    phantom/ + examples/ made by __graft_entry__.build(); it travels to the GPU box) through
    the third-party stub shim.  None if the copy is absent."""
    root = REF_COPY if os.path.isdir(os.path.join(REF_COPY, "phantom")) else None
    if root is None and os.path.isdir("/root/reference/phantom"):
        root = "/root/reference"
    if root is None:
        return None, None
    os.environ["PHX_REFERENCE_ROOT"] = root
    from oracle import ref_shim

    ref_shim.REFERENCE_ROOT = root
    return ref_shim.import_reference(), ref_shim


def _cpu_env(kind: str, cfg: str, worker: int):
    """(env, step_fn(t), T): one env object of config `cfg` on the reference (`kind` =
    "reference") or on the oracle port, with its pre-generated action tape."""
    r = np.random.RandomState(1000 + worker)
    if kind == "reference":
        ph, shim = _reference_module()
        assert ph is not None, "baseline/_ref is absent"
    else:
        import oracle.phantom_oracle as ph
    if cfg == "C2":
        T = 100
        acts = r.uniform(0, 100, size=(T, 1)).astype(np.float32)
        if kind == "reference":  # the reference's own example file, imported unmodified
            env = shim.import_reference_supply_chain().SupplyChainEnv()
        else:
            from oracle.workloads import supply_chain as wl
            env = wl.build(ph, _NativeStream())
        return env, (lambda t: env.step({"SHOP": acts[t]})), T
    if cfg == "C3":
        from oracle.workloads import market as wl
        T = 99
        env = wl.build(ph, _NativeStream(), num_steps=T)
        ids = env.strategic_agent_ids
        a = r.uniform(0, 1, size=(T, len(ids), 1)).astype(np.float32)
        tape = [{aid: (a[t, s] if s < wl.N_MAKERS else int(a[t, s, 0] > 0.4))
                 for s, aid in enumerate(ids)} for t in range(T)]
    elif cfg == "C4":
        from oracle.workloads import stackelberg as wl
        T = 100
        env = wl.build(ph, _NativeStream(), num_steps=T)
        ids = env.strategic_agent_ids
        a = r.uniform(0, 1, size=(T, len(ids), 1)).astype(np.float32)
        tape = [{aid: a[t, s] for s, aid in enumerate(ids)} for t in range(T)]
    elif cfg == "C5":
        from oracle.workloads import dense as wl
        T = 8
        env = wl.build(ph, num_steps=T)
        ids = env.strategic_agent_ids
        a = r.uniform(0, 1, size=(T, len(ids), 1)).astype(np.float32)
        tape = [{aid: a[t, s] for s, aid in enumerate(ids)} for t in range(T)]
    else:
        raise ValueError(cfg)
    return env, (lambda t: env.step(tape[t])), T


def _cpu_worker(args):
    """Steps ONE env object for `seconds` (whole episodes incl. reset) and returns
    (env-steps, elapsed).  One process = one core."""
    kind, cfg, worker, seconds, warm_steps = args
    os.environ.setdefault("PYTHONHASHSEED", "1")
    env, step, T = _cpu_env(kind, cfg, worker)
    env.reset()
    for t in range(min(warm_steps, T)):
        step(t)
    env.reset()
    n, t0 = 0, time.perf_counter()
    while True:
        for t in range(T):
            step(t)
        n += T
        env.reset()
        if time.perf_counter() - t0 >= seconds:
            break
    return n, time.perf_counter() - t0


def cpu_rate(kind: str, cfg: str, seconds: float, cores: int, warm_steps: int = 20):
    """Aggregate env-steps/s of `cores` processes, each stepping its own env object."""
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        t0 = time.perf_counter()
        res = pool.map(_cpu_worker, [(kind, cfg, w, seconds, warm_steps) for w in range(cores)])
        wall = time.perf_counter() - t0
    return sum(n / dt for n, dt in res), sum(n for n, _ in res), wall


def cpu_kind():
    """"reference" when the unmodified reference's sources are present (baseline/_ref, or
    /root/reference in the build container); the import itself happens in the worker processes."""
    return ("reference" if os.path.isdir(os.path.join(REF_COPY, "phantom")) or
            os.path.isdir("/root/reference/phantom") else "port")


def cpu_baseline(cfg: str, seconds: float, cores: int, kind=None):
    kind = kind or cpu_kind()
    rate, steps, wall = cpu_rate(kind, cfg, seconds, cores)
    what = ("the UNMODIFIED reference (baseline/_ref, imported through oracle/ref_shim.py; "
            "native np.random)" if kind == "reference" else
            "oracle.phantom_oracle (Python restatement of PhantomEnv.step; baseline/_ref absent)")
    out = {"value": rate, "unit": "env-steps/s", "cores": cores, "kind": kind,
           "sample": (f"{cores} processes x {seconds:.0f} s, one env object each, whole "
                      f"{CONFIGS[cfg]['T']}-step episodes incl. reset, on {what}; {steps} "
                      f"env-steps, wall {wall:.1f} s"),
           "cpu_model": cpu_model()}
    if kind == "port":
        out["port_over_reference"] = port_calibration()
    return out


def cpu_single_core(cfg: str, kind: str, windows: int = 5, seconds: float = 2.0):
    """one env object, one core, median of `windows` timed windows after a
    200-step warm-up."""
    ctx = mp.get_context("spawn")
    with ctx.Pool(1) as pool:
        res = [pool.apply(_cpu_worker, ((kind, cfg, 0, seconds, 200),)) for _ in range(windows)]
    rates = sorted(n / dt for n, dt in res)
    return {"value": rates[len(rates) // 2], "unit": "env-steps/s", "cores": 1, "kind": kind,
            "windows": windows, "min": rates[0], "max": rates[-1]}


# ----------------------------------------------------------------------------- our arm
def _timed_launches(torch, dev, launch, K: int, W: int, world: int):
    """W warm-up launches, then EXACTLY K launches timed with CUDA events on the launch stream.
    The K launches are captured into CUDA graphs of up to 1000 nodes (launch i uses buffer set
    i % NBUF, baked into the node), so the timed region holds no host launch gaps; falls back
    to eager launches if the capture is refused."""
    import torch.distributed as dist

    for i in range(W):
        launch(i, torch.cuda.current_stream(dev).cuda_stream)
    torch.cuda.synchronize()
    glen = min(K, 1000)
    graph, mode = None, "eager"
    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side, capture_error_mode="relaxed"):
            for i in range(glen):
                launch(W + i, torch.cuda.current_stream(dev).cuda_stream)
        g.replay()
        torch.cuda.synchronize()
        graph, mode = g, f"cuda-graph({glen} launches per replay)"
    except Exception:
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    done = 0
    if graph is not None:
        for _ in range(K // glen):
            graph.replay()
        done = (K // glen) * glen
    for i in range(done, K):
        launch(W + i, torch.cuda.current_stream(dev).cuda_stream)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms, mode


def run_config(name, args, torch, dev, rank, world, local):
    """Device-timed launches + e2e (host buffers) of one config on this rank's shard.  Returns
    the sub-record pieces (rank 0 assembles)."""
    import torch.distributed as dist

    from phantom_b200 import _lib as L

    c = CONFIGS[name]
    T, S, O = c["T"], c["S"], c["O"]
    if c["scaling"] == "weak":
        E, offset = c["E"], rank * c["E"]
    else:
        E, offset = c["E"] // world, rank * (c["E"] // world)
    env = make_env(name, num_envs=E, seed=SEED, device=local, env_offset=offset, auto_reset=True)
    env.reset_batch()
    if c.get("specialise"):
        try:
            env.specialise()
        except Exception as exc:  # no nvcc on the box: the generic kernel still measures
            print(f"[bench] {name}: specialise() failed, generic kernel: {exc!r}", file=sys.stderr)
    NBUF = c["nbuf"]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts, outs = [], []
    for _ in range(NBUF):
        a = torch.rand((T, E, S, 1), generator=gen, device=dev)
        if c["binary_from"] is not None:
            a[:, :, c["binary_from"]:] = (a[:, :, c["binary_from"]:] > 0.4).float()
        if c["act_scale"] != 1.0:
            a *= c["act_scale"]
        acts.append(a)
        outs.append(env._alloc_outputs((T,)))
    lean = c["lean"]

    def launch(i, stream):
        o = outs[i % NBUF]
        p = lambda t: None if lean else t.data_ptr()
        L.check(L.lib.phx_rollout(env._handle, T, acts[i % NBUF].data_ptr(), None,
                                  o.observations.data_ptr(), p(o.obs_mask), o.rewards.data_ptr(),
                                  p(o.reward_mask), p(o.terminations), p(o.truncations),
                                  o.all_done.data_ptr(), stream))

    K = args.steps if name == "C2" else max(3, min(args.steps, args.sub_steps))
    W = args.warmup if name == "C2" else max(3, min(args.warmup, 5))
    ms, mode = _timed_launches(torch, dev, launch, K, W, world)
    env.check_errors()
    rec = {"E": E, "T": T, "K": K, "W": W, "ms": ms, "mode": mode, "kernel": env.exec_name,
           "env": env, "acts": acts, "outs": outs}
    if args.timed_only:
        return rec

    # ---- e2e: the same metric through the host-buffer C-ABI call (pinned host memory, H2D of
    # the actions and D2H of every output plane the device-timed launch writes, inside the call)
    e2e_steps = max(3, min(K, 12)) if name == "C2" else 3
    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    h_act = pin((T, E, S, 1), torch.float32)
    h_act.copy_(acts[0].cpu())
    h_obs, h_rew = pin((T, E, S, O), torch.float32), pin((T, E, S), torch.float32)
    h_all = pin((T, E, 2), torch.uint8)
    h_u8 = [None] * 4 if lean else [pin((T, E, S), torch.uint8) for _ in range(4)]
    hp = lambda t: None if t is None else t.data_ptr()

    def launch_host():
        L.check(L.lib.phx_rollout_host(env._handle, T, h_act.data_ptr(), None, h_obs.data_ptr(),
                                       hp(h_u8[0]), h_rew.data_ptr(), hp(h_u8[1]), hp(h_u8[2]),
                                       hp(h_u8[3]), h_all.data_ptr()))

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
    rec.update(e2e_s=e2e_s, e2e_steps=e2e_steps, h2d=h_act.numel() * 4,
               d2h=(h_obs.numel() * 4 + h_rew.numel() * 4 + h_all.numel() +
                    sum(t.numel() for t in h_u8 if t is not None)))
    del h_act, h_obs, h_rew, h_all, h_u8
    return rec


def config_record(name, rec, world, peak, peak_src, with_cpu, args):
    c = CONFIGS[name]
    E, T, K = rec["E"], rec["T"], rec["K"]
    launch_ms = rec["ms"] / K
    value = world * E * T * K / (rec["ms"] * 1e-3)
    bytes_per_launch = E * (T * c["b_io"] + c["b_state"])
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    out = {
        "workload": c["workload"], "kernel": rec["kernel"], "envs_per_gpu": E,
        "transitions_per_launch": T, "scaling": c["scaling"], "steps": K, "warmup": rec["W"],
        "value": value, "unit": "env-steps/s", "ms_per_step": launch_ms, "timing": rec["mode"],
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "algorithmic_bytes_per_env_step": c["b_io"] + c["b_state"] / T,
                     "peak_source": peak_src},
    }
    if name != "C2" and world == 1 and E == c["E"]:  # (a capture of the full-size one-GPU launch)
        out["roofline"]["traffic"], out["roofline"]["traffic_source"] = ncu_traffic(name)
    if "e2e_s" in rec:
        e2e_value = world * E * T * rec["e2e_steps"] / rec["e2e_s"]
        out["e2e"] = {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": rec["h2d"],
                      "d2h_bytes_per_step": rec["d2h"], "steps": rec["e2e_steps"],
                      "api": "phx_rollout_host (pinned host buffers)"}
    if with_cpu:
        out["cpu_baseline"] = cpu_baseline(name, args.cpu_seconds if name == "C2" else
                                           args.sub_cpu_seconds, os.cpu_count() or 1)
    return out


def bench_gather(torch, dev, env, acts, rank, world, T, E):
    """The trainer gather in the path: every launch writes its lean planes
    straight into a packed block (sharding.PackedOutputs) and ONE NCCL collective per launch
    moves it, double-buffered so that the collective of launch i overlaps the kernel of launch
    i+1.  Timed on the device, max over ranks."""
    import torch.distributed as dist

    from phantom_b200.sharding import LEAN_PLANES, PackedOutputs, gather_packed

    total = world * E
    res = {}
    for label, lead_T, dst in (("rollout_T100_allgather", T, None), ("rollout_T100_gather0", T, 0),
                               ("step_T1_allgather", None, None)):
        packed = [PackedOutputs.for_env(env, lead_T, LEAN_PLANES, total_envs=total,
                                        world_size=world) for _ in range(2)]
        recv = [torch.empty((world, packed[0].nbytes), dtype=torch.uint8, device=dev)
                if (dst is None or rank == dst) else None for _ in range(2)]
        n_launch = 10 if lead_T else 100
        pending = [None, None]

        def one(i):
            b = i % 2
            if pending[b] is not None:
                pending[b].wait()
            a = acts[i % len(acts)]
            packed[b].launch(env, a if lead_T else a[i % T])
            g = gather_packed(packed[b], total, dst=dst, out=recv[b], async_op=True)
            pending[b] = g

        for i in range(3):
            one(i)
        for p in pending:
            if p is not None:
                p.wait()
        pending[:] = [None, None]
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n_launch):
            one(i)
        for p in pending:
            if p is not None:
                p.wait()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        steps_per_launch = lead_T or 1
        nbytes = packed[0].nbytes
        # bytes ENTERING the busiest GPU per launch: (world - 1) blocks
        rx = (world - 1) * nbytes
        res[label] = {
            "us_per_launch": ms * 1e3 / n_launch, "launches": n_launch,
            "block_bytes_per_rank": nbytes,
            "value_with_gather": world * E * steps_per_launch * n_launch / (ms * 1e-3),
            "unit": "env-steps/s", "rx_bytes_per_launch": rx,
            "rx_GBps": rx * n_launch / (ms * 1e-3) / 1e9,
            "nvlink_peak_GBps": 770.0,  # measured peer-copy rate per direction (B200_PROFILING.md)
            "collective": "ncclAllGather" if dst is None else "ncclGather(dst=0) = grouped send/recv",
        }
        env.check_errors()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

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
    # host threads of the e2e path (compact-wire expansion inside phx_rollout_host): an equal
    # share of the cores this rank may run on
    os.environ.setdefault("PHX_HOST_THREADS",
                          str(max(1, min(32, len(os.sched_getaffinity(0)) // max(world, 1)))))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from phantom_b200 import _lib as L

    names = [n for n in args.configs.split(",") if n in CONFIGS]
    if "C2" not in names:
        names.insert(0, "C2")
    recs = {}
    with ClockSampler(local) as clocks:
        head = run_config("C2", args, torch, dev, rank, world, local)
    recs["C2"] = head
    env, acts, outs = head["env"], head["acts"], head["outs"]
    E, T = head["E"], head["T"]

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
        c2 = CONFIGS["C2"]
        single = {"api": "phx_step, T=1, 100 launches per CUDA graph", "us_per_launch": us,
                  "value": E / (us * 1e-6), "unit": "env-steps/s",
                  "bytes_per_launch": E * (c2["b_io"] + c2["b_state"])}
        env.check_errors()
        env.reset_batch()

    # ---- the same launches with ALL SEVEN output planes (obs_mask / reward_mask / term / trunc
    # are constant for this env class, which is why the headline passes NULL for them)
    full_io = None
    if not args.timed_only:
        def launch_full(i, stream):
            o = outs[i % len(outs)]
            L.check(L.lib.phx_rollout(env._handle, T, acts[i % len(acts)].data_ptr(), None,
                                      o.observations.data_ptr(), o.obs_mask.data_ptr(),
                                      o.rewards.data_ptr(), o.reward_mask.data_ptr(),
                                      o.terminations.data_ptr(), o.truncations.data_ptr(),
                                      o.all_done.data_ptr(), stream))

        kf = max(3, min(head["K"], 200))
        ms_full, _ = _timed_launches(torch, dev, launch_full, kf, 5, world)
        env.check_errors()
        c2c = CONFIGS["C2"]
        bytes_full = E * (T * (c2c["b_io"] + 4) + c2c["b_state"])
        full_io = {"ms_per_step": ms_full / kf, "steps": kf,
                   "value": world * E * T * kf / (ms_full * 1e-3), "unit": "env-steps/s",
                   "algorithmic_bytes_per_launch": bytes_full,
                   "achieved_GBps": bytes_full / (ms_full / kf * 1e-3) / 1e9,
                   "planes": "obs, obs_mask, reward, reward_mask, term, trunc, all_done"}

    if args.timed_only:  # profiling aid: only the device-timed region (tools/profile_round.sh)
        if rank == 0:
            print(json.dumps({"timed_only": True, "ms_per_step": head["ms"] / head["K"],
                              "gpu_launches": head["K"]}), flush=True)
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return

    gather = None
    if world > 1:
        try:
            gather = bench_gather(torch, dev, env, acts, rank, world, T, E)
        except Exception as exc:  # never lose the headline to the secondary measurement
            gather = {"error": repr(exc)}
    env.close()
    del head["env"], head["acts"], head["outs"], env, acts, outs
    torch.cuda.empty_cache()

    sub_clocks = {}
    for name in names:
        if name == "C2":
            continue
        try:
            with ClockSampler(local) as ck:
                r = run_config(name, args, torch, dev, rank, world, local)
            r["env"].close()
            del r["env"], r["acts"], r["outs"]
            torch.cuda.empty_cache()
            recs[name] = r
            sub_clocks[name] = ck.summary()
        except Exception as exc:
            recs[name] = {"error": repr(exc)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        traffic, traffic_src = ncu_traffic()
        with_cpu = world == 1 and not args.no_cpu_baseline
        c2 = config_record("C2", head, world, peak, peak_src, with_cpu, args)
        nb = CONFIGS["C2"]["nbuf"]
        line = {
            "metric": "env-steps/sec (65k parallel supply-chain envs)",
            "value": c2["value"], "unit": "env-steps/s", "n_gpus": world, "steps": head["K"],
            "warmup": head["W"], "ms_per_step": c2["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {
                "workload": c2["workload"],
                "envs_per_gpu": E, "transitions_per_launch": T, "kernel": c2["kernel"],
                "l2": f"inputs larger than L2: {nb} rotating action/output sets of "
                      f"{(E * T * 4 + c2['roofline']['algorithmic_bytes_per_launch']) / 1e6:.0f} MB",
                "timing": c2["timing"],
                "parallelism": f"env-sharded x{world}, no data-path collective",
            },
            "roofline": dict(c2["roofline"], traffic=traffic, traffic_source=traffic_src,
                             kernel="sc_fast2_kernel<0,1,0> (lean I/O, vectorised action ring; programmatic dependent launch)"),
            "e2e": dict(c2["e2e"], bound="pcie / host expansion",
                        d2h_GBps_per_gpu=c2["e2e"]["d2h_bytes_per_step"] * c2["e2e"]["value"] /
                        (world * E * T) / 1e9,
                        host_threads=int(os.environ.get("PHX_HOST_THREADS", "1")),
                        **pcie_ceiling(world, c2["e2e"]["value"])),
            "single_step": single,
            "full_io": None if full_io is None else dict(
                full_io, frac=full_io["achieved_GBps"] / peak, peak=peak),
            # kernels of ours inside the device-timed region (one per bench step); the e2e region
            # launches one kernel per pipeline chunk (8 per call)
            "gpu_launches": head["K"],
            "gpu_launches_e2e": head["e2e_steps"] * 8,
            "clocks": clocks.summary(),
        }
        if "cpu_baseline" in c2:
            line["cpu_baseline"] = c2["cpu_baseline"]
        if gather is not None:
            line["gather"] = gather
        subs = {}
        for name in names:
            if name == "C2":
                continue
            r = recs[name]
            if "error" in r:
                subs[name] = r
                continue
            subs[name] = config_record(name, r, world, peak, peak_src, with_cpu, args)
            subs[name]["clocks"] = sub_clocks.get(name)
            line["gpu_launches"] += r["K"]
        line["configs"] = subs
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------- reference arm
def run_reference(args):
    """The reference's own CPU implementation of the path on the box's host cores: the
    UNMODIFIED reference from baseline/_ref when present (kind "reference"), else the oracle
    port.  Each "step" = one bounded all-core sample of cpu_seconds."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    kind = cpu_kind()
    t0 = time.perf_counter()
    if args.warmup:
        cpu_rate(kind, "C2", min(2.0, args.cpu_seconds), cores)
    samples = [cpu_baseline("C2", args.cpu_seconds, cores, kind)
               for _ in range(max(1, min(args.steps, 3)))]
    best = max(samples, key=lambda r: r["value"])
    value = float(np.mean([r["value"] for r in samples]))
    single = cpu_single_core("C2", kind)
    subs = {}
    for name in [n for n in args.configs.split(",") if n in CONFIGS and n != "C2"]:
        try:
            subs[name] = {"workload": CONFIGS[name]["workload"],
                          "cpu_baseline": cpu_baseline(name, args.sub_cpu_seconds, cores, kind)}
            subs[name]["value"] = subs[name]["cpu_baseline"]["value"]
        except Exception as exc:
            subs[name] = {"error": repr(exc)}
    line = {
        "impl": "reference", "metric": "env-steps/sec (65k parallel supply-chain envs)",
        "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": len(samples),
        "warmup": 1 if args.warmup else 0, "ms_per_step": args.cpu_seconds * 1e3,
        "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "python-int", "data": "synthetic",
        "config": {"workload": CONFIGS["C2"]["workload"],
                   "cpu": f"same dynamics on the host ({kind}): one env object per process, "
                          f"{cores} processes, 100-step episodes incl. reset; each step = a "
                          f"{args.cpu_seconds:.0f} s sample; {cpu_model()}"},
        "cpu_baseline": dict(best, value=value),
        "single_core": single,
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "configs": subs,
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--configs", default="C2,C3,C4,C5")
    ap.add_argument("--sub-steps", type=int, default=10,
                    help="timed launches of the C3/C4/C5 sub-records (<= --steps)")
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    ap.add_argument("--sub-cpu-seconds", type=float, default=4.0)
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
