#!/usr/bin/env python
"""Block-engine timing: digital_ads_market at its shipped size (40 + 40 + 40 advertisers,
122 agents) and the same env at 32 agents on the tile engine, T-step auto-reset rollouts.

    python tools/bench_wide.py [--envs 8192] [--steps 20] [--reps 10] [--once]
"""
import argparse
import json
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(per_theme, E, mode):
    from phantom_b200.envs import digital_ads_market as da
    from phantom_b200.utils.samplers import UniformFloatSampler

    budgets = ([(5.0, 15.001, 5.0, 15.0)] * per_theme + [(7.0, 17.001, 7.0, 17.0)] * per_theme +
               [(10.0, 20.001, 10.0, 20.0)] * per_theme)
    st = {f"ADV_{i + 1}": da.AdvertiserAgent.Supertype(budget=UniformFloatSampler(*b))
          for i, b in enumerate(budgets)}
    env = da.DigitalAdsEnv(num_steps=20, num_agents_theme={"travel": per_theme, "tech": per_theme, "sport": per_theme},
                           strategy="first", agent_supertypes=st, exec_mode=mode, num_envs=E, seed=1,
                           auto_reset=True)
    env.reset_batch()
    return env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true", help="one launch per case (for ncu)")
    a = ap.parse_args()
    out = {}
    for name, per_theme, mode in (("wide_122_agents", 40, "auto"), ("wide_32_agents", 10, "wide"),
                                  ("tile_32_agents", 10, "queue")):
        env = build(per_theme, a.envs, mode)
        S = 3 * per_theme
        A = torch.rand(a.steps, a.envs, S, 1, device="cuda") * 0.5
        env.rollout_batch(A)
        torch.cuda.synchronize()
        if a.once:
            env.close()
            continue
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(a.reps):
            env.rollout_batch(A)
        ev1.record()
        torch.cuda.synchronize()
        env.check_errors()
        ms = ev0.elapsed_time(ev1) / a.reps
        out[name] = {"exec": env.exec_name, "agents": S + 2, "envs": a.envs, "T": a.steps,
                     "ms_per_launch": round(ms, 4),
                     "env_steps_per_s": round(a.envs * a.steps / ms * 1e3),
                     "agent_steps_per_s": round(a.envs * a.steps * (S + 2) / ms * 1e3)}
        env.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
