#!/usr/bin/env python
"""Secondary measurements: env-steps/s and HBM-roofline fraction of every BASELINE config on
one GPU (C2 on both kernels, C3, C4, C5).  One JSON line per config.  bench.py (the driver's
contract) stays the C2 headline; this script feeds DESIGN.md's per-kernel table.

    python tools/bench_configs.py [--steps K] [--only C3]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from phantom_b200 import _lib as L  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def measure(name, make_env, S, O, T, E, b_state, steps, binary_from=None, lean=False):
    env = make_env(num_envs=E, seed=0, auto_reset=True)
    env.reset_batch()
    if name.endswith("-jit"):
        env.specialise()  # step kernel rebuilt with this env class as a compile-time constant
    dev = torch.device("cuda", 0)
    nbuf = max(2, int(np.ceil(200e6 / (T * E * S * (4 + 4 * O + 8)))))
    gen = torch.Generator(device=dev).manual_seed(7)
    acts, outs = [], []
    for _ in range(nbuf):
        a = torch.rand((T, E, S, 1), generator=gen, device=dev)
        if binary_from is not None:
            a[:, :, binary_from:] = (a[:, :, binary_from:] > 0.4).float()
        if name.startswith("C2"):
            a *= 100.0
        acts.append(a)
        outs.append(env._alloc_outputs((T,)))
    stream = torch.cuda.current_stream(dev).cuda_stream

    def launch(i):
        o = outs[i % nbuf]
        p = lambda t: None if (lean and t is not o.observations and t is not o.rewards
                               and t is not o.all_done) else t.data_ptr()
        L.check(L.lib.phx_rollout(env._handle, T, acts[i % nbuf].data_ptr(), None,
                                  p(o.observations), p(o.obs_mask), p(o.rewards), p(o.reward_mask),
                                  p(o.terminations), p(o.truncations), p(o.all_done), stream))

    for i in range(3):
        launch(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        launch(i)
    e1.record()
    torch.cuda.synchronize()
    env.check_errors()
    ms = e0.elapsed_time(e1) / steps
    b_io = S * (4 + 4 * O + 4 + (0 if lean else 4)) + 2
    bytes_per_launch = E * (T * b_io + b_state)
    gbs = bytes_per_launch / (ms * 1e-3) / 1e9
    line = {"config": name, "kernel": env.exec_name, "envs": E, "T": T, "agents_strategic": S,
            "ms_per_launch": ms, "env_steps_per_s": E * T / (ms * 1e-3),
            "algorithmic_bytes_per_env_step": b_io + b_state / T, "achieved_GBps": gbs,
            "frac_of_measured_hbm_peak": gbs / peak()}
    print(json.dumps(line), flush=True)
    env.close()


def _ads_env(per_theme, **k):
    from phantom_b200.envs import digital_ads_market as da
    from phantom_b200.utils.samplers import UniformFloatSampler as U

    b = ([(5.0, 15.001, 5.0, 15.0)] * per_theme + [(7.0, 17.001, 7.0, 17.0)] * per_theme +
         [(10.0, 20.001, 10.0, 20.0)] * per_theme)
    st = {f"ADV_{i + 1}": da.AdvertiserAgent.Supertype(budget=U(*x)) for i, x in enumerate(b)}
    return da.DigitalAdsEnv(num_steps=20, num_agents_theme={"travel": per_theme, "tech": per_theme,
                                                            "sport": per_theme},
                            agent_supertypes=st, **k)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from phantom_b200.envs.dense import DenseEnv
    from phantom_b200.envs.market import MarketEnv
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv
    from phantom_b200.envs.supply_chain import SupplyChainEnv
    from phantom_b200.envs import simple_market as sm

    cfgs = [
        ("C2-fast", lambda **k: SupplyChainEnv(exec_mode="fast", **k), 1, 3, 100, 65536, 48, None, True),
        ("C2-queue", lambda **k: SupplyChainEnv(exec_mode="queue", **k), 1, 3, 100, 65536,
         2 * (16 + 8 + 8 * 16), None, False),
        ("C3-market", lambda **k: MarketEnv(**k), 31, 3, 99, 32768,
         2 * (16 + 8 + 32 * (8 * 4 + 4 + 12) + 8), 7, False),
        ("C2-thread", lambda **k: SupplyChainEnv(exec_mode="thread", **k), 1, 3, 100, 65536,
         2 * (16 + 8 + 8 * 16), None, False),
        ("C4-stackelberg-queue", lambda **k: StackelbergGameEnv(exec_mode="queue", **k), 4, 2, 100, 131072,
         2 * (16 + 8 + 8 * (16 + 4) + 4), None, False),
        ("C4-stackelberg-thread", lambda **k: StackelbergGameEnv(exec_mode="thread", **k), 4, 2, 100, 131072,
         2 * (16 + 8 + 8 * (16 + 4) + 4), None, False),
        ("C2-thread-jit", lambda **k: SupplyChainEnv(exec_mode="thread", **k), 1, 3, 100, 65536,
         2 * (16 + 8 + 8 * 16), None, False),
        ("C4-stackelberg-thread-jit", lambda **k: StackelbergGameEnv(exec_mode="thread", **k), 4, 2, 100,
         131072, 2 * (16 + 8 + 8 * (16 + 4) + 4), None, False),
        ("C3-market-jit", lambda **k: MarketEnv(**k), 31, 3, 99, 32768,
         2 * (16 + 8 + 32 * (8 * 4 + 4 + 12) + 8), 7, False),
        ("C4-stackelberg-queue-jit", lambda **k: StackelbergGameEnv(exec_mode="queue", **k), 4, 2, 100,
         131072, 2 * (16 + 8 + 8 * (16 + 4) + 4), None, False),
        ("C5-dense", lambda **k: DenseEnv(**k), 128, 3, 8, 16384, 2 * (16 + 6 * 128 * 4), None, False),
        # the reference's simple_market example (3 buyers + 2 sellers, 10-step episodes); state =
        # header + 19 words x 8 slots + reward / obs caches + env words
        ("X-simple-market-thread", lambda **k: sm.example_env(num_steps=10, exec_mode="thread", **k),
         5, 3, 100, 65536, 2 * (16 + 8 + 19 * 8 * 4 + 8 * 4 + 8 * 12 + 8 + 8), 0, False),
        ("X-simple-market-thread-jit", lambda **k: sm.example_env(num_steps=10, exec_mode="thread", **k),
         5, 3, 100, 65536, 2 * (16 + 8 + 19 * 8 * 4 + 8 * 4 + 8 * 12 + 8 + 8), 0, False),
        # a 32-agent market (7 sellers + 25 buyers; SURVEY 6 probed the reference at ~1.5 k
        # env-steps/s per core on 8 + 24): lane-per-agent tiling, compact acting queue
        ("X-simple-market-32", lambda **k: sm.example_env(
            tuple((0.5, 0.1, 0.9) for _ in range(25)), 7, 10, **k),
         32, 3, 50, 32768, 2 * (16 + 8 + 19 * 32 * 4 + 32 * 4 + 32 * 12 + 8 + 8), 0, False),
        ("X-simple-market-32-jit", lambda **k: sm.example_env(
            tuple((0.5, 0.1, 0.9) for _ in range(25)), 7, 10, **k),
         32, 3, 50, 32768, 2 * (16 + 8 + 19 * 32 * 4 + 32 * 4 + 32 * 12 + 8 + 8), 0, False),
        # the reference's digital_ads_market example, 10 + 10 + 10 advertisers (SURVEY 6 probed the
        # reference at ~650 env-steps/s per core on 122 agents)
        ("X-digital-ads-32", lambda **k: _ads_env(10, **k), 30, 3, 40, 32768,
         2 * (16 + 8 + 15 * 32 * 4 + 32 * 4 + 32 * 12 + 8 + 32 * 4), None, False),
        ("X-digital-ads-8-thread", lambda **k: _ads_env(2, exec_mode="thread", **k), 6, 3, 100, 65536,
         2 * (16 + 8 + 15 * 8 * 4 + 8 * 4 + 8 * 12 + 8 + 8 * 4), None, False),
        ("X-digital-ads-8-thread-jit", lambda **k: _ads_env(2, exec_mode="thread", **k), 6, 3, 100, 65536,
         2 * (16 + 8 + 15 * 8 * 4 + 8 * 4 + 8 * 12 + 8 + 8 * 4), None, False),
        ("X-simple-market-queue", lambda **k: sm.example_env(num_steps=10, exec_mode="queue", **k),
         5, 3, 100, 65536, 2 * (16 + 8 + 19 * 8 * 4 + 8 * 4 + 8 * 12 + 8 + 8), 0, False),
    ]
    for name, mk, S, O, T, E, bst, binf, lean in cfgs:
        if args.only and args.only not in name:
            continue
        measure(name, mk, S, O, T, E, bst, args.steps, binf, lean)


if __name__ == "__main__":
    main()
