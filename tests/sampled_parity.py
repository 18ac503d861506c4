"""Full-size parity by sampling: a launch at the BASELINE size of a config
is compared, bit for bit, with the object-level oracle run on a random sample of GLOBAL env ids.
The RNG contract is keyed by the global env index, so env e of a 32 768-env launch and a
one-env oracle built with `StepStream(seed, e, ...)` must agree on every output plane of every
step and on the final agent state.

The oracle envs are independent, so the sample is spread over a process pool (the dense env
costs ~0.15 s of Python per step: 16 256 messages).  Workers only import `oracle/` (CPU)."""
from __future__ import annotations

import multiprocessing as mp
import os

import numpy as np


def _market_state(env):
    """[31, 5] like oracle.make_golden.market_state / tests.test_gpu_market.market_state."""
    rows = []
    for a in env.agents.values():
        n = type(a).__name__
        if n == "MakerAgent":
            rows.append([a.inventory, a.cash, a.last_price, a.last_notional, 0])
        elif n == "TakerAgent":
            rows.append([a.value, a.best_price, a.best_maker, a.holdings, a.last_surplus])
    return np.array(rows, np.int64)


def _oracle_episode(args):
    """Worker: one oracle episode of env `e` of config `cfg` on its action tape [T,S,1]."""
    cfg, seed, e, actions = args
    import oracle.phantom_oracle as po
    from oracle import harness, rng

    T, S = actions.shape[:2]
    mask = np.ones((1, T, S), np.uint8)
    if cfg == "market":
        from oracle.workloads import market as wl

        st = rng.StepStream(seed, e, wl.STREAM_TAKER_VALUE)
        env = wl.build(po, st, num_steps=T)
        tr = harness.run_generic(env, harness.EpisodeClock([st]), actions[None], mask, 3,
                                 state_fn=_market_state)
    elif cfg == "stackelberg":
        from oracle.workloads import stackelberg as wl

        st = rng.StepStream(seed, e, wl.STREAM_FOLLOWER_VALUE)
        env = wl.build(po, st, num_steps=T)
        tr = harness.run_generic(env, harness.EpisodeClock([st]), actions[None], mask, 2,
                                 state_fn=wl.state)
    elif cfg == "dense":
        from oracle.workloads import dense as wl

        env = wl.build(po, n_agents=S, num_steps=T)
        tr = harness.run_generic(env, harness.EpisodeClock([]), actions[None], mask, 3,
                                 state_fn=wl.state)
    else:  # pragma: no cover
        raise ValueError(cfg)
    return e, {k: v[0] for k, v in tr.items() if k != "messages"}


def oracle_sample(cfg: str, seed: int, env_ids, actions_TES1: np.ndarray, workers=None):
    """{env id: trace} for the sampled env ids; `actions_TES1` holds only the sampled envs'
    tapes, [T, n_sample, S, 1] in the order of `env_ids`."""
    jobs = [(cfg, seed, int(e), np.ascontiguousarray(actions_TES1[:, i]))
            for i, e in enumerate(env_ids)]
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2) - 1, 32))
    if workers == 1:
        return dict(_oracle_episode(j) for j in jobs)
    with mp.get_context("spawn").Pool(workers) as pool:
        return dict(pool.map(_oracle_episode, jobs, chunksize=max(1, len(jobs) // (4 * workers))))


def assert_rollout_rows_equal_oracle(out, env_ids, traces, offset: int = 0):
    """Rows `env_ids - offset` of a device rollout (BatchStep with a leading T axis) == the
    oracle traces: masks / flags bit-exact, obs exact float32, rewards == float32(float64)."""
    idx = np.asarray(env_ids) - offset
    import torch

    sel = torch.as_tensor(idx, device=out.observations.device)
    obs = out.observations[:, sel].cpu().numpy()
    om = out.obs_mask[:, sel].cpu().numpy()
    rew = out.rewards[:, sel].cpu().numpy()
    rm = out.reward_mask[:, sel].cpu().numpy()
    te = out.terminations[:, sel].cpu().numpy()
    tr = out.truncations[:, sel].cpu().numpy()
    ad = out.all_done[:, sel].cpu().numpy()
    for i, e in enumerate(env_ids):
        g = traces[int(e)]
        assert np.array_equal(om[:, i], g["obs_mask"]), (e, "obs_mask")
        m = g["obs_mask"].astype(bool)
        assert np.array_equal(obs[:, i][m], g["obs"][m]), (e, "obs")
        assert np.array_equal(rm[:, i], g["reward_mask"]), (e, "reward_mask")
        m = g["reward_mask"] == 1
        assert np.array_equal(rew[:, i][m], g["reward"].astype(np.float32)[m]), (e, "reward")
        assert np.array_equal(te[:, i], g["term"]), (e, "term")
        assert np.array_equal(tr[:, i], g["trunc"]), (e, "trunc")
        assert np.array_equal(ad[:, i], g["all_done"]), (e, "all_done")


def sample_ids(E: int, n: int, seed: int, offset: int = 0):
    r = np.random.RandomState(seed)
    ids = np.sort(r.choice(E, size=min(n, E), replace=False))
    # always include the first and the last env of the launch (block / tile edges)
    ids[0], ids[-1] = 0, E - 1
    return ids + offset
