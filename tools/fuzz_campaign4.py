#!/usr/bin/env python
"""One-off campaign 4 (GPU box): random CONFIGURATIONS of the authored BASELINE workloads, inside
the domains their device programs state (a first run with wider ranges only met the loud
refusals "market family: up to 7 makers and 24 takers" / "stackelberg family: up to 7
followers") -- C3 market (1-7 makers, 1-24 takers, shuffled batches or not), C4 Stackelberg
game (1-7 followers), C5 dense graph (2-128 agents, random symmetric graphs of density 1 / 0.5 / 0.1,
round limits) -- with missing actions, two episodes each, on every tiling / kernel that takes
them, against the oracle port's run of the same reference-API workload: every output plane of
every step and the agent state.

    python tools/fuzz_campaign4.py [--count 40]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from tests.generic_parity import run_device_vs_golden  # noqa: E402
from tests.sampled_parity import _market_state  # noqa: E402
from tests.test_gpu_dense import dense_state  # noqa: E402
from tests.test_gpu_market import market_state  # noqa: E402
from tests.test_gpu_stackelberg import game_state  # noqa: E402


def traces_to_golden(per_env, A, M, seed):
    g = {k: np.stack([t[k] for t in per_env]) for k in per_env[0] if k != "messages"}
    g.update(actions=A, action_mask=M, seed=np.int64(seed))
    return g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=40)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle import harness, rng
    from oracle.make_golden import generic_actions
    from oracle.workloads import dense as wd
    from oracle.workloads import market as wm
    from oracle.workloads import stackelberg as ws
    from phantom_b200.envs.dense import DenseEnv
    from phantom_b200.envs.market import MarketEnv
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    runs, bad = 0, []

    def check(tag, cfg, modes, make, g, state_fn, **kw):
        nonlocal runs
        for mode in modes:
            runs += 1
            try:
                run_device_vs_golden(lambda **k: make(mode, **k), g, state_fn, **kw).close()
            except Exception as exc:
                bad.append((tag, cfg, mode))
                print("MISMATCH", tag, cfg, mode, type(exc).__name__, str(exc)[:300], flush=True)

    for c in range(a.count):
        r = np.random.RandomState(60000 + c)
        n_env, n_ep = 3, 2
        # ---- C3 market
        nm = int(r.randint(1, 8))
        nt = int(r.randint(1, 25))
        T, seed, shuffle = int(r.randint(4, 25)), 70000 + c, bool(r.uniform() < 0.3)
        A, M = generic_actions(n_env, n_ep, T, nm + nt, discrete_from=nm, seed=c, p_missing=0.1)
        per_env = []
        for e in range(n_env):
            st = rng.StepStream(seed, e, wm.STREAM_TAKER_VALUE)
            env = wm.build(po, st, n_makers=nm, n_takers=nt, num_steps=T, shuffle_batches=shuffle)
            clock = harness.EpisodeClock([st])
            slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
            with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
                per_env.append(harness.run_generic(env, clock, A[e], M[e], 3, state_fn=_market_state))
        modes = ["queue"]  # (the market program has no thread-per-env form)
        check("market", (nm, nt, T, shuffle), modes,
              lambda mode, **k: MarketEnv(nm, nt, num_steps=T, shuffle_batches=shuffle, exec_mode=mode, **k),
              traces_to_golden(per_env, A, M, seed), market_state)
        # ---- C4 Stackelberg game
        nf = int(r.randint(1, 8))
        T, seed = int(r.randint(4, 31)), 80000 + c
        A, M = generic_actions(n_env, n_ep, T, nf + 1, seed=c + 1, p_missing=0.1)
        per_env = []
        for e in range(n_env):
            st = rng.StepStream(seed, e, ws.STREAM_FOLLOWER_VALUE)
            env = ws.build(po, st, n_followers=nf, num_steps=T)
            per_env.append(harness.run_generic(env, harness.EpisodeClock([st]), A[e], M[e], 2, state_fn=ws.state))
        modes = ["queue"] + (["thread"] if nf + 1 <= 8 else [])
        check("stackelberg", (nf, T), modes,
              lambda mode, **k: StackelbergGameEnv(nf, num_steps=T, exec_mode=mode, **k),
              traces_to_golden(per_env, A, M, seed), game_state)
        # ---- C5 dense graph
        n = int(r.choice([2, 3, 5, 8, 12, 17, 31, 32, 33, 64, 100, 127, 128]))
        dens = float(r.choice([1.0, 0.5, 0.1]))
        up = np.triu((r.uniform(size=(n, n)) < dens).astype(np.int64), 1)
        adj = up + up.T
        T, seed = int(r.randint(2, 7)), 90000 + c
        rl = [2, 2, 3, None][int(r.randint(4))]
        A, M = generic_actions(n_env, n_ep, T, n, seed=c + 2, p_missing=0.1 if r.uniform() < 0.5 else 0.0)
        per_env = []
        for e in range(n_env):
            env = wd.build(po, n_agents=n, adjacency=adj, num_steps=T, round_limit=rl)
            per_env.append(harness.run_generic(env, harness.EpisodeClock([]), A[e], M[e], 3, state_fn=wd.state))
        check("dense", (n, dens, T, rl), ["auto"],
              lambda mode, **k: DenseEnv(n, adj, num_steps=T, round_limit=rl, **k),
              traces_to_golden(per_env, A, M, seed), dense_state)
    print(json.dumps({"device_runs": runs, "mismatches": bad}))


if __name__ == "__main__":
    main()
