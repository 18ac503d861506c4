#!/usr/bin/env python
"""One-off campaign 3 (GPU box): (a) tests/test_gpu_fuzz.py's random supply-chain topologies for
many more seeds; (b) random PhantomEnv / StackelbergEnv env classes over the mock agents --
halving and request/response echo agents, random graphs, ignore_connection_errors, round limits
(RuntimeError), sends without an edge (NetworkError), agents that terminate mid-episode, leader /
follower lists in random order, message tracking -- on every tiling against the CPU oracle port:
observations, rewards, done flags, call counters, float32 levels, exception types AND the
tracked message list of every step.

    python tools/fuzz_campaign3.py [--first 0] [--count 1000] [--chains 300]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from tests.kat_scenarios import counts  # noqa: E402  (before anything can shadow `tests`)


def random_mock_env(K, case_seed, **kw):
    r = np.random.RandomState(50000 + case_seed)
    ph = K.ph
    strat = [f"s{i}" for i in range(int(r.randint(1, 4)))]
    echo = [f"e{i}" for i in range(int(r.randint(1, 5)))]
    agents = [K.MockStrategicAgent(a, num_steps=(int(r.randint(1, 7)) if r.uniform() < 0.3 else None))
              for a in strat]
    agents += [K.EchoAgent(e, seed_value=int(r.choice([0, 0, 3, 4, 9, 17])),
                           request_response=bool(r.uniform() < 0.3)) for e in echo]
    agents = [agents[i] for i in r.permutation(len(agents))]
    round_limit = None if r.uniform() < 0.6 else int(r.randint(1, 5))
    shuffle = bool(r.uniform() < 0.35)  # BatchResolver(shuffle_batches=True): the contract's Fisher-Yates
    network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=True, round_limit=round_limit,
                                                            shuffle_batches=shuffle),
                         ignore_connection_errors=bool(r.uniform() < 0.3))
    for i in range(len(echo)):
        for j in range(i + 1, len(echo)):
            if r.uniform() < 0.6:
                network.add_connection(echo[i], echo[j])
    if r.uniform() < 0.15:  # an echo agent next to a strategic one: no handler there (ValueError)
        network.add_connection(echo[0], strat[0])
    kind = "stackelberg" if r.uniform() < 0.5 else "base"
    net = K.finish_network(network)
    if kind == "base":
        env = ph.PhantomEnv(num_steps=8, network=net, **kw)
    else:
        everyone = list(r.permutation(strat + echo))
        cut = int(r.randint(1, len(everyone)))
        env = ph.StackelbergEnv(8, net, [str(x) for x in everyone[:cut]], [str(x) for x in everyone[cut:]], **kw)
    return env, strat, echo, shuffle


def run_mock_env(K, case_seed):
    import contextlib

    seed = 77 + case_seed
    is_device = hasattr(K.ph.PhantomEnv, "default_exec_mode")
    env, strat, echo, shuffle = random_mock_env(K, case_seed, **({"seed": seed} if is_device else {}))
    clock, patch = None, contextlib.nullcontext()
    if not is_device:  # the oracle / reference: np.random.shuffle -> the contract's shuffle
        from oracle import harness

        clock = harness.EpisodeClock([])
        slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
        patch = harness.patched_np_shuffle(seed, 0, clock, env, slot_of)

    def plain(d):
        return {k: (None if v is None else
                    [round(float(x), 6) for x in np.asarray(v, np.float64).reshape(-1)])
                for k, v in d.items()}

    def msgs():
        out = []
        for m in env.network.resolver.tracked_messages:
            p = m.payload
            out.append([str(m.sender_id), str(m.receiver_id), type(p).__name__,
                        int(getattr(p, "value", getattr(p, "cash", 0)))])
        return out

    trace = [("shuffle", shuffle)]
    try:
      with patch:
        if clock is not None:
            clock.on_reset()
        obs, _ = env.reset()
        trace.append(("reset", plain(obs)))
        for t in range(8):
            env.network.resolver.clear_tracked_messages()
            if clock is not None:
                clock.on_step(env)
            step = env.step({a: np.array([0]) for a in strat})
            trace.append((
                "step", plain(step.observations), plain(step.rewards),
                {k: bool(v) for k, v in step.terminations.items()},
                {k: bool(v) for k, v in step.truncations.items()},
                [list(map(int, counts(env.agents[a]))) for a in strat],
                [[int(env.agents[e].handled_count), int(env.agents[e].handled_total),
                  float(env.agents[e].level)] for e in echo], msgs()))
            if step.terminations["__all__"] or step.truncations["__all__"]:
                break
    except Exception as exc:
        trace.append(("raise", type(exc).__name__))
    if hasattr(env, "close"):
        env.close()
    return trace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=1000)
    ap.add_argument("--chains", type=int, default=300)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle.workloads import mock as omock
    from tools.fuzz_campaign import device_ns

    KO, KD = omock.build_classes(po), device_ns()
    bad = []
    for s in range(a.first, a.first + a.count):
        want = json.loads(json.dumps(run_mock_env(KO, s)))
        # (shuffled batches need the per-receiver lists of the tile / block engines)
        for mode in (("queue", "wide") if want[0][1] else ("thread", "queue", "wide")):
            KD.ph.PhantomEnv.default_exec_mode = mode
            try:
                got = json.loads(json.dumps(run_mock_env(KD, s)))
            except Exception as exc:
                got = ["exception", type(exc).__name__, str(exc)[:200]]
            if got != want:
                bad.append((s, mode))
                k = next((i for i, (x, y) in enumerate(zip(got, want)) if x != y), -1)
                print("MISMATCH mock", s, mode, "first differing record", k, str(got[k] if 0 <= k < len(got) else got)[:400],
                      "WANT", str(want[k] if 0 <= k < len(want) else want)[:400], flush=True)
    KD.ph.PhantomEnv.default_exec_mode = "auto"
    print("mock", {"cases": a.count, "mismatches": bad}, flush=True)
    chain_bad = []
    import tests.test_gpu_fuzz as tf

    for s in range(100, 100 + a.chains):
        for mode in ("thread", "queue"):
            try:
                tf.test_random_supply_chain_topology(s, mode)
            except Exception as exc:
                chain_bad.append((s, mode))
                print("MISMATCH chain", s, mode, type(exc).__name__, str(exc)[:300], flush=True)
    print(json.dumps({"mock_cases": a.count, "mock_mismatches": bad, "chain_cases": a.chains,
                      "chain_mismatches": chain_bad}))


if __name__ == "__main__":
    main()
