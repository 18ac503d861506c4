"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Drives envs built with the reference plugin API (real reference or phantom_oracle) under
the counter-based RNG contract and records tensor traces in the layout the C-ABI uses
([env, strategic_agent, ...]), so that tests can compare them with the CUDA path.
"""
from __future__ import annotations

import contextlib
from typing import Any, Callable, Dict, List, Optional, Sequence

import numpy as np

from . import rng


class EpisodeClock:
    """Keeps the (episode, step) coordinates of every RNG stream of one env in sync with
    the env's own clock."""

    def __init__(self, streams: Sequence[rng.StepStream]):
        self.streams = list(streams)
        self.episode = -1

    def on_reset(self) -> None:
        self.episode += 1
        for s in self.streams:
            s.begin(self.episode, 0)

    def on_step(self, env) -> None:
        for s in self.streams:
            s.begin(self.episode, env.current_step + 1)


@contextlib.contextmanager
def patched_np_randint(stream: rng.StepStream):
    """Route `np.random.randint(n)` (the reference example's call site,
    supply_chain.py:64) to the contract stream while the block runs."""
    orig = np.random.randint
    np.random.randint = lambda n, *a, **k: stream.randint(n)
    try:
        yield
    finally:
        np.random.randint = orig


@contextlib.contextmanager
def patched_np_uniform(stream: rng.StepStream):
    """Route `np.random.uniform(low, high)` (the reference's UniformFloatSampler.sample,
    phantom/utils/samplers.py:142) to the contract stream while the block runs."""
    orig = np.random.uniform
    np.random.uniform = lambda low=0.0, high=1.0, size=None: stream.uniform(low, high)
    try:
        yield
    finally:
        np.random.uniform = orig


@contextlib.contextmanager
def patched_np_random(stream: rng.StepStream):
    """Route `np.random.random()` (StochasticNetwork.add_connection / resample_connectivity,
    phantom/network.py:389,446) to the contract stream: the c-th call after the stream's
    begin() -- i.e. the c-th base connection of this reset -- gets draw idx c as a float64
    d24 * 2^-24."""
    orig = np.random.random
    np.random.random = lambda size=None: stream.next_d24() / 16777216.0
    try:
        yield
    finally:
        np.random.random = orig


def contract_shuffle(seed: int, env: int, episode: int, step: int, recv_slot: int, k: int,
                     items: list) -> None:
    """The contract's replacement of `np.random.shuffle(batch)` (phantom/resolvers.py:151):
    Fisher-Yates from the back, as numpy's legacy shuffle walks it --
        for i = n-1 .. 1:  j = randint(i + 1);  swap(items[i], items[j])
    with randint(m) = (d24 * m) >> 24 and the draw taken from
        stream = 0x100 + receiver slot,  idx = k * 256 + (n - 1 - i)
    where k counts the batches this receiver has handled in this env step (one per round in
    which it had mail).  Keyed by receiver so that all receivers of a round can shuffle in
    parallel on the device."""
    n = len(items)
    assert n <= 256, "contract_shuffle: batches of at most 256 messages"
    for i in range(n - 1, 0, -1):
        d = rng.d24(seed, env, episode, step, 0x100 + recv_slot, k * 256 + (n - 1 - i))
        j = rng.randint(i + 1, d)
        items[i], items[j] = items[j], items[i]


@contextlib.contextmanager
def patched_np_shuffle(seed: int, env_index: int, clock: "EpisodeClock", env, slot_of):
    """Route `np.random.shuffle(batch)` inside BatchResolver.resolve to contract_shuffle.
    The receiver is read off the batch; the per-receiver batch counter restarts whenever the
    env's (episode, step) coordinate changes."""
    orig = np.random.shuffle
    seen = {"key": None, "count": {}}

    def shuffle(batch):
        key = (clock.episode, env.current_step)
        if seen["key"] != key:
            seen["key"], seen["count"] = key, {}
        if len(batch) == 0:
            return
        r = slot_of[batch[0].receiver_id]
        k = seen["count"].get(r, 0)
        seen["count"][r] = k + 1
        contract_shuffle(seed, env_index, clock.episode, env.current_step, r, k, batch)

    np.random.shuffle = shuffle
    try:
        yield
    finally:
        np.random.shuffle = orig


def tracked_to_rows(tracked, slot_of: Dict[Any, int], type_of: Callable[[Any], int],
                    value_of: Callable[[Any], Sequence[float]]) -> np.ndarray:
    """Flatten Resolver.tracked_messages to rows (sender_slot, recv_slot, type, v0, v1)."""
    rows = []
    for m in tracked:
        v = list(value_of(m.payload)) + [0.0, 0.0]
        rows.append((slot_of[m.sender_id], slot_of[m.receiver_id], type_of(m.payload), v[0], v[1]))
    return np.asarray(rows, dtype=np.float64).reshape(-1, 5)


def run_supply_chain(env, clock: EpisodeClock, actions: np.ndarray,
                     action_mask: Optional[np.ndarray] = None,
                     stream_ctx=contextlib.nullcontext, track: bool = False) -> Dict[str, np.ndarray]:
    """Run `actions.shape[0]` episodes of `actions.shape[1]` steps on one supply-chain env.

    actions: f32 [n_ep, T, 1]; action_mask: u8 [n_ep, T] (0 => SHOP absent from `actions`,
    which makes env.py:330-333 fall back to generate_messages()).
    """
    from .workloads import supply_chain as wl

    n_ep, T = actions.shape[:2]
    out = {
        "reset_obs": np.zeros((n_ep, 3), np.float32),
        "obs": np.zeros((n_ep, T, 3), np.float32),
        "reward": np.zeros((n_ep, T), np.float64),
        "term": np.zeros((n_ep, T), np.uint8),
        "trunc": np.zeros((n_ep, T), np.uint8),
        "all_term": np.zeros((n_ep, T), np.uint8),
        "all_trunc": np.zeros((n_ep, T), np.uint8),
        "state": np.zeros((n_ep, T, 4), np.int64),
    }
    msgs: List[np.ndarray] = []
    slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
    with stream_ctx():
        for ep in range(n_ep):
            clock.on_reset()
            obs, _ = env.reset()
            out["reset_obs"][ep] = obs["SHOP"]
            for t in range(T):
                clock.on_step(env)
                acts = {}
                if action_mask is None or action_mask[ep, t]:
                    acts["SHOP"] = actions[ep, t]
                if track:
                    env.network.resolver.clear_tracked_messages()
                step = env.step(acts)
                out["obs"][ep, t] = step.observations["SHOP"]
                out["reward"][ep, t] = step.rewards["SHOP"]
                out["term"][ep, t] = step.terminations["SHOP"]
                out["trunc"][ep, t] = step.truncations["SHOP"]
                out["all_term"][ep, t] = step.terminations["__all__"]
                out["all_trunc"][ep, t] = step.truncations["__all__"]
                out["state"][ep, t] = wl.shop_state(env)
                if track:
                    rows = tracked_to_rows(
                        env.network.resolver.tracked_messages, slot_of,
                        lambda p: wl.PAYLOAD_TYPE_IDS[type(p).__name__],
                        lambda p: (p.size,))
                    hdr = np.tile(np.array([[ep, t]], np.float64), (len(rows), 1))
                    msgs.append(np.concatenate([hdr, rows], axis=1))
    if track:
        out["messages"] = np.concatenate(msgs, axis=0) if msgs else np.zeros((0, 7))
    return out


def run_generic(env, clock: EpisodeClock, actions: np.ndarray, action_mask: np.ndarray,
                obs_dim: int, track: bool = False, state_fn=None,
                convert=None, flatten=None) -> Dict[str, np.ndarray]:
    """Run actions.shape[0] episodes x actions.shape[1] steps of ANY env built with the
    plugin API and record the tensors of the C-ABI layout.

    actions f32 [n_ep, T, S, A]; action_mask u8 [n_ep, T, S] (0 => agent absent from the
    `actions` mapping).  Discrete action spaces receive int(round(a[0])); `convert(s, a)`
    overrides how strategic agent s's action row becomes the value handed to the env.
    Outputs (S = strategic agents in env order):
      obs f32 [n_ep,T,S,O] (zero padded), obs_mask u8, reward f64, reward_mask u8 (0 absent,
      1 value, 2 None), term/trunc u8 (255 = key absent), all_done u8 [..,2];
      reset_obs / reset_mask for every episode; optional state [n_ep,T,...] via state_fn(env).
    """
    n_ep, T, S, A = actions.shape
    ids = env.strategic_agent_ids
    assert len(ids) == S
    discrete = [hasattr(getattr(env.agents[a], "action_space", None), "n") for a in ids]
    out = {
        "reset_obs": np.zeros((n_ep, S, obs_dim), np.float32),
        "reset_mask": np.zeros((n_ep, S), np.uint8),
        "obs": np.zeros((n_ep, T, S, obs_dim), np.float32),
        "obs_mask": np.zeros((n_ep, T, S), np.uint8),
        "reward": np.zeros((n_ep, T, S), np.float64),
        "reward_mask": np.zeros((n_ep, T, S), np.uint8),
        "term": np.full((n_ep, T, S), 255, np.uint8),
        "trunc": np.full((n_ep, T, S), 255, np.uint8),
        "all_done": np.zeros((n_ep, T, 2), np.uint8),
    }
    states, msgs = [], []
    slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}

    def put_obs(dst, dmask, obs):
        for s, aid in enumerate(ids):
            if aid in obs:
                v = (flatten(obs[aid]) if flatten is not None
                     else np.asarray(obs[aid], np.float32).reshape(-1))
                dst[s, : v.size] = v
                dmask[s] = 1

    for ep in range(n_ep):
        clock.on_reset()
        obs, _ = env.reset()
        put_obs(out["reset_obs"][ep], out["reset_mask"][ep], obs)
        ep_states = []
        for t in range(T):
            clock.on_step(env)
            acts = {}
            for s, aid in enumerate(ids):
                if action_mask[ep, t, s]:
                    a = actions[ep, t, s]
                    if convert is not None:
                        acts[aid] = convert(s, a)
                    else:
                        acts[aid] = int(round(float(a[0]))) if discrete[s] else a
            if track:
                env.network.resolver.clear_tracked_messages()
            step = env.step(acts)
            put_obs(out["obs"][ep, t], out["obs_mask"][ep, t], step.observations)
            for s, aid in enumerate(ids):
                if aid in step.rewards:
                    r = step.rewards[aid]
                    out["reward_mask"][ep, t, s] = 2 if r is None else 1
                    out["reward"][ep, t, s] = 0.0 if r is None else r
                if aid in step.terminations:
                    out["term"][ep, t, s] = step.terminations[aid]
                    out["trunc"][ep, t, s] = step.truncations[aid]
            out["all_done"][ep, t] = (step.terminations["__all__"], step.truncations["__all__"])
            if state_fn is not None:
                ep_states.append(state_fn(env))
            if track:
                for k, m in enumerate(env.network.resolver.tracked_messages):
                    conv = float if track == "raw" else int  # "raw": keep float payload fields
                    vals = [conv(v) for v in m.payload.__dict__.values()][:2] + [0, 0]
                    msgs.append((ep, t, slot_of[m.sender_id], slot_of[m.receiver_id],
                                 type(m.payload).__name__, vals[0], vals[1]))
        if state_fn is not None:
            states.append(np.stack(ep_states))
    if state_fn is not None:
        out["state"] = np.stack(states)
    if track:
        out["messages"] = msgs
    return out
