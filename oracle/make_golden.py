"""TEST INFRASTRUCTURE.  Generates tests/golden/*.npz from the UNMODIFIED reference.

Only runs in the build container (needs /root/reference, imported through
oracle/ref_shim.py).  The committed fixtures are what the oracle restatement and the CUDA
path are compared against where the reference is not available (GPU box).

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
import sys

import numpy as np

from . import harness, ref_shim, rng

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")


def supply_chain_actions(n_env: int, n_ep: int, T: int, seed: int = 123):
    """Actions covering the quirks of SURVEY.md 7.2-H6: in-range uniforms, exact
    integers, exact halves (round-half-even), values above max stock, negatives,
    tiny values, and steps where the shop supplies no action at all."""
    r = np.random.RandomState(seed)
    a = r.uniform(0, 100, size=(n_env, n_ep, T, 1)).astype(np.float32)
    mask = np.ones((n_env, n_ep, T), np.uint8)
    kinds = r.randint(0, 12, size=(n_env, n_ep, T))
    a[kinds == 0] = np.floor(a[kinds == 0])                      # exact integers
    a[kinds == 1] = np.floor(a[kinds == 1]) + np.float32(0.5)    # ties -> even
    a[kinds == 2] = a[kinds == 2] * np.float32(3.0)              # above max stock
    if n_env > 4:
        a[4:][kinds[4:] == 3] *= np.float32(-0.25)               # negative requests
        mask[5:][kinds[5:] == 4] = 0                             # action missing
    a[kinds == 5] *= np.float32(1e-3)
    return a, mask


def gen_supply_chain_reference() -> None:
    """The reference's own example file, RNG call site patched to the contract."""
    sc = ref_shim.import_reference_supply_chain()
    seed, n_env, n_ep, T = 20261017, 12, 2, 100
    actions, mask = supply_chain_actions(n_env, n_ep, T)
    keys = None
    per_env = []
    for e in range(n_env):
        stream = rng.PackedStream(seed, e, 0, 5, 5)  # supply_chain.py:12,64: 5 customers, randint(5)
        env = sc.SupplyChainEnv()
        env.network.resolver.enable_tracking = e < 3
        clock = harness.EpisodeClock([stream])
        tr = harness.run_supply_chain(
            env, clock, actions[e], mask[e],
            stream_ctx=lambda s=stream: harness.patched_np_randint(s), track=e < 3)
        per_env.append(tr)
        keys = keys or [k for k in tr if k != "messages"]
    out = {k: np.stack([t[k] for t in per_env]) for k in keys}
    rows = []
    for e in range(3):
        m = per_env[e]["messages"]
        rows.append(np.concatenate([np.full((len(m), 1), e, np.float64), m], axis=1))
    out["messages"] = np.concatenate(rows)  # (env, ep, t, sender, recv, type, v0, v1)
    out["actions"], out["action_mask"] = actions, mask
    out["seed"] = np.int64(seed)
    np.savez_compressed(os.path.join(GOLDEN, "supply_chain_reference.npz"), **out)
    print("supply_chain_reference.npz:", {k: v.shape for k, v in out.items()})


def generic_actions(n_env, n_ep, T, S, discrete_from=None, seed=11, p_missing=0.05):
    r = np.random.RandomState(seed)
    a = r.uniform(0, 1, size=(n_env, n_ep, T, S, 1)).astype(np.float32)
    if discrete_from is not None:
        a[:, :, :, discrete_from:, :] = (a[:, :, :, discrete_from:, :] > 0.3)
    m = (r.uniform(size=(n_env, n_ep, T, S)) > p_missing).astype(np.uint8)
    return a, m


MESSAGE_TYPE_IDS = {"Quote": 0, "Order": 1, "Fill": 2}


def pack_generic(per_env, actions, mask, seed, msg_envs, type_ids):
    keys = [k for k in per_env[0] if k != "messages"]
    out = {k: np.stack([t[k] for t in per_env]) for k in keys}
    rows = []
    for e in range(msg_envs):
        for (ep, t, s, r, name, v0, v1) in per_env[e]["messages"]:
            rows.append((e, ep, t, s, r, type_ids[name], v0, v1))
    out["messages"] = np.asarray(rows, np.int64).reshape(-1, 8)
    out["actions"], out["action_mask"], out["seed"] = actions, mask, np.int64(seed)
    return out


def market_state(env):
    rows = []
    for a in env.agents.values():
        n = type(a).__name__
        if n == "MakerAgent":
            rows.append([a.inventory, a.cash, a.last_price, a.last_notional, 0])
        elif n == "TakerAgent":
            rows.append([a.value, a.best_price, a.best_maker, a.holdings, a.last_surplus])
    return np.array(rows, np.int64)


def gen_market_reference() -> None:
    """oracle/workloads/market.py (C3) executed by the UNMODIFIED reference."""
    from .workloads import market

    ref = ref_shim.import_reference()
    seed, n_env, n_ep, T, S = 20261018, 6, 2, 99, market.N_MAKERS + market.N_TAKERS
    actions, mask = generic_actions(n_env, n_ep, T, S, discrete_from=market.N_MAKERS)
    per_env = []
    for e in range(n_env):
        st = rng.StepStream(seed, e, market.STREAM_TAKER_VALUE)
        env = market.build(ref, st, enable_tracking=e < 1)
        per_env.append(harness.run_generic(env, harness.EpisodeClock([st]), actions[e], mask[e], 3,
                                           track=e < 1, state_fn=market_state))
        if e >= 1:
            per_env[-1]["messages"] = []
    out = pack_generic(per_env, actions, mask, seed, 1, MESSAGE_TYPE_IDS)
    np.savez_compressed(os.path.join(GOLDEN, "market_reference.npz"), **out)
    print("market_reference.npz:", {k: v.shape for k, v in out.items()},
          "terminated:", int((out["term"] == 1).sum()))


def gen_stackelberg_reference() -> None:
    """oracle/workloads/stackelberg.py (C4) executed by the UNMODIFIED reference."""
    from .workloads import stackelberg as wl

    ref = ref_shim.import_reference()
    seed, n_env, n_ep, T, S = 20261019, 16, 2, 100, 1 + wl.N_FOLLOWERS
    actions, mask = generic_actions(n_env, n_ep, T, S, seed=13, p_missing=0.1)
    per_env = []
    for e in range(n_env):
        st = rng.StepStream(seed, e, wl.STREAM_FOLLOWER_VALUE)
        env = wl.build(ref, st, enable_tracking=e < 4)
        per_env.append(harness.run_generic(env, harness.EpisodeClock([st]), actions[e], mask[e], 2,
                                           track=e < 4, state_fn=wl.state))
        if e >= 4:
            per_env[-1]["messages"] = []
    out = pack_generic(per_env, actions, mask, seed, 4, wl.MESSAGE_TYPE_IDS)
    np.savez_compressed(os.path.join(GOLDEN, "stackelberg_reference.npz"), **out)
    print("stackelberg_reference.npz:", {k: v.shape for k, v in out.items()})


def gen_dense_reference() -> None:
    """oracle/workloads/dense.py (C5) executed by the UNMODIFIED reference: the full 128-agent
    complete graph (no message list: 16 256 per step) and a sparse 12-agent graph with the
    tracked message order."""
    from .workloads import dense as wl

    ref = ref_shim.import_reference()
    for name, n, adj, n_env, T, msg_envs in (
            ("dense128_reference.npz", 128, None, 2, 8, 0),
            ("dense12_reference.npz", 12, wl.random_adjacency(12, 0.6, 1), 6, 8, 6)):
        seed, n_ep = 20261020, 2
        actions, mask = generic_actions(n_env, n_ep, T, n, seed=17, p_missing=0.12)
        per_env = []
        for e in range(n_env):
            env = wl.build(ref, n_agents=n, adjacency=adj, num_steps=T, enable_tracking=msg_envs > 0)
            per_env.append(harness.run_generic(env, harness.EpisodeClock([]), actions[e], mask[e], 3,
                                               track=msg_envs > 0, state_fn=wl.state))
            if msg_envs == 0:
                per_env[-1]["messages"] = []
        out = pack_generic(per_env, actions, mask, seed, msg_envs, wl.MESSAGE_TYPE_IDS)
        if adj is not None:
            out["adjacency"] = adj
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, {k: v.shape for k, v in out.items()})


def gen_supply_chain2_reference() -> None:
    """oracle/workloads/supply_chain2.py (tutorial-2 env with supertypes) on the UNMODIFIED
    reference, its UniformFloatSampler drawing through the patched np.random.uniform."""
    from .workloads import supply_chain2 as wl

    ref = ref_shim.import_reference()
    seed, n_env, n_ep, T, S = 20261021, 10, 2, 100, wl.N_SHOPS
    r = np.random.RandomState(19)
    actions = r.uniform(0, 100, size=(n_env, n_ep, T, S, 1)).astype(np.float32)
    mask = (r.uniform(size=(n_env, n_ep, T, S)) > 0.08).astype(np.uint8)
    per_env, weights = [], []
    for e in range(n_env):
        streams = {s: rng.StepStream(seed, e, s)
                   for s in (wl.STREAM_ORDER, wl.STREAM_SAMPLER, wl.STREAM_SHOP_CHOICE)}
        ws = []
        with harness.patched_np_uniform(streams[wl.STREAM_SAMPLER]):
            env = wl.build(ref, streams, ref.utils.samplers.UniformFloatSampler,
                           enable_tracking=e < 3)
            per_env.append(harness.run_generic(
                env, harness.EpisodeClock(list(streams.values())), actions[e], mask[e], 4,
                track=e < 3, state_fn=lambda env: (ws.append(wl.weights(env)), wl.state(env))[1]))
        if e >= 3:
            per_env[-1]["messages"] = []
        weights.append(np.array(ws).reshape(n_ep, T, S))
    out = pack_generic(per_env, actions, mask, seed, 3, wl.MESSAGE_TYPE_IDS)
    out["weights"] = np.stack(weights)
    np.savez_compressed(os.path.join(GOLDEN, "supply_chain2_reference.npz"), **out)
    print("supply_chain2_reference.npz:", {k: v.shape for k, v in out.items()})


def gen_shuffle_and_stochastic_reference() -> None:
    """SURVEY 8(f) rows 2 and 4 on the UNMODIFIED reference:
    market_shuffle_reference.npz       C3 market with BatchResolver(shuffle_batches=True)
    supply_chain2_stochastic_reference.npz   tutorial-2 env on a StochasticNetwork
                                       (ignore_connection_errors) with shuffled batches;
                                       records every episode's resampled graph."""
    from .workloads import market
    from .workloads import supply_chain2 as wl

    ref = ref_shim.import_reference()
    # ---- market + shuffle
    seed, n_env, n_ep, T, S = 20261022, 4, 2, 99, market.N_MAKERS + market.N_TAKERS
    actions, mask = generic_actions(n_env, n_ep, T, S, discrete_from=market.N_MAKERS, seed=23)
    per_env = []
    for e in range(n_env):
        st = rng.StepStream(seed, e, market.STREAM_TAKER_VALUE)
        env = market.build(ref, st, enable_tracking=e < 1, shuffle_batches=True)
        clock = harness.EpisodeClock([st])
        slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
        with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
            per_env.append(harness.run_generic(env, clock, actions[e], mask[e], 3,
                                               track=e < 1, state_fn=market_state))
        if e >= 1:
            per_env[-1]["messages"] = []
    out = pack_generic(per_env, actions, mask, seed, 1, MESSAGE_TYPE_IDS)
    np.savez_compressed(os.path.join(GOLDEN, "market_shuffle_reference.npz"), **out)
    print("market_shuffle_reference.npz:", {k: v.shape for k, v in out.items()})

    # ---- supply chain 2 on a stochastic network, with and without shuffled batches
    for shuffle, name in ((True, "supply_chain2_stochastic_reference.npz"),
                          (False, "supply_chain2_stochastic_plain_reference.npz")):
        seed, n_env, n_ep, T, S = 20261023, 12, 3, 40, wl.N_SHOPS
        rates = (0.75, 0.625)
        r = np.random.RandomState(29)
        actions = r.uniform(0, 100, size=(n_env, n_ep, T, S, 1)).astype(np.float32)
        mask = (r.uniform(size=(n_env, n_ep, T, S)) > 0.08).astype(np.uint8)
        per_env, adjs = [], []
        for e in range(n_env):
            streams = {s: rng.StepStream(seed, e, s)
                       for s in (wl.STREAM_ORDER, wl.STREAM_SAMPLER, wl.STREAM_SHOP_CHOICE,
                                 wl.STREAM_CONNECTIVITY)}
            adj = []
            with harness.patched_np_uniform(streams[wl.STREAM_SAMPLER]), \
                    harness.patched_np_random(streams[wl.STREAM_CONNECTIVITY]):
                env = wl.build(ref, streams, ref.utils.samplers.UniformFloatSampler, num_steps=T,
                               enable_tracking=e < 4, rates=rates, shuffle_batches=shuffle)
                clock = harness.EpisodeClock(list(streams.values()))
                slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
                with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
                    per_env.append(harness.run_generic(
                        env, clock, actions[e], mask[e], 4, track=e < 4,
                        state_fn=lambda env: (adj.append(wl.adjacency(env)), wl.state(env))[1]))
            if e >= 4:
                per_env[-1]["messages"] = []
            adjs.append(np.array(adj).reshape(n_ep, T, 8, 8)[:, 0])  # the graph of each episode
        out = pack_generic(per_env, actions, mask, seed, 4, wl.MESSAGE_TYPE_IDS)
        out["adjacency"] = np.stack(adjs)
        out["rates"] = np.array(rates)
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, {k: v.shape for k, v in out.items()}, "mean degree", out["adjacency"].mean())


SIMPLE_MARKET_WIDE_BUYERS = ((0.3, 0.1, 0.9), (0.8, 0.5, 1.0), (0.6, 0.0, 1.0), (0.5, 0.25, 0.25))


def gen_simple_market_reference(only=slice(None)) -> None:
    """The reference's own examples/environments/simple_market modules, UNMODIFIED (imported by
    oracle/workloads/simple_market.py:build_reference), under the contract RNG:
      simple_market_reference.npz       the example script's cast: 3 buyers + 2 sellers, 10 steps
      simple_market_wide_reference.npz  4 buyers with non-degenerate type samplers + 3 sellers
                                        (np.mean over three prices), 12 steps
      simple_market_9s_reference.npz    3 buyers + 9 sellers: np.mean enters numpy's pairwise
                                        summation (8 accumulators) -- 12 agents, 16-lane tiles
      simple_market_block_reference.npz 28 buyers + 12 sellers = 40 agents: the 128-lane block
                                        engine; pairwise sum of 8 + 3 leftover prices"""
    from .workloads import simple_market as wl

    rb = np.random.RandomState(20261029)
    block_buyers = tuple((float(np.round(p, 3)), float(np.round(lo, 3)), float(np.round(lo + w, 3)))
                         for p, lo, w in zip(rb.uniform(0.2, 0.9, 28), rb.uniform(0.0, 0.6, 28),
                                             rb.uniform(0.0, 0.4, 28)))
    for name, buyers, n_sellers, T, n_env, n_ep, seed in (
            ("simple_market_reference.npz", wl.EXAMPLE_BUYERS, wl.EXAMPLE_SELLERS, 10, 8, 3, 20261024),
            ("simple_market_wide_reference.npz", SIMPLE_MARKET_WIDE_BUYERS, 3, 12, 6, 2, 20261025),
            ("simple_market_9s_reference.npz", SIMPLE_MARKET_WIDE_BUYERS[:3], 9, 10, 4, 2, 20261030),
            ("simple_market_block_reference.npz", block_buyers, 12, 8, 3, 2, 20261031))[only]:
        actions, mask = wl.actions_for(n_env, n_ep, T, len(buyers), n_sellers, seed % 1000)
        per_env = []
        for e in range(n_env):
            coords = wl.Coords(seed, e)
            with wl.contract_rng(coords, {f"b{i + 1}": i for i in range(len(buyers))}):
                env, _ = wl.build_reference(buyers, n_sellers, T)
                clock = harness.EpisodeClock([coords])
                env.network.resolver.enable_tracking = e < 2
                tr = harness.run_generic(env, clock, actions[e], mask[e], wl.OBS_DIM,
                                         state_fn=wl.state, convert=wl.to_action(env),
                                         track="raw" if e < 2 else False)
            per_env.append(tr)
        rows = [(e, ep, t, s, r, {"Price": 0, "Order": 1}[name], v0)
                for e in range(2) for (ep, t, s, r, name, v0, v1) in per_env[e]["messages"]]
        for tr in per_env:
            tr["messages"] = []
        out = pack_generic(per_env, actions, mask, seed, 0, {})
        # Resolver.tracked_messages of envs 0 and 1: (env, episode, step, sender slot, receiver
        # slot, type, value) with the float64 price / the order volume as the value
        out["messages"] = np.asarray(rows, np.float64).reshape(-1, 7)
        out["buyers"] = np.array(buyers, np.float64)
        out["n_sellers"] = np.int64(n_sellers)
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, {k: v.shape for k, v in out.items()})


# the tape of tests/test_gpu_simple_market.py::test_simple_market_handler_driven_on_env_word
SIMPLE_MARKET_HANDLER_TAPE = dict(T=10, n_env=12, n_ep=2, seed=5, action_seed=9)


def gen_simple_market_handler_reference() -> None:
    """simple_market_handler_reference.npz: the reference's simple_market example classes,
    UNMODIFIED, with an env HANDLER on the Sellers stage (fsm.py:294-307;
    workloads/simple_market.py:seller_handler_avg_price -- resolve_network(), then Buyers once
    avg_price >= 0.5, else Sellers again), under the contract RNG.  Same cast and action tape as
    the device-vs-oracle GPU test, so  device == oracle (GPU test)  and  oracle == reference
    (tests/test_oracle_golden.py, this fixture)  close the chain."""
    from .workloads import simple_market as wl

    c = SIMPLE_MARKET_HANDLER_TAPE
    buyers, n_sellers = wl.EXAMPLE_BUYERS, wl.EXAMPLE_SELLERS
    actions, mask = wl.actions_for(c["n_env"], c["n_ep"], c["T"], len(buyers), n_sellers,
                                   c["action_seed"])
    per_env = []
    for e in range(c["n_env"]):
        coords = wl.Coords(c["seed"], e)
        with wl.contract_rng(coords, {f"b{i + 1}": i for i in range(len(buyers))}):
            env, _ = wl.build_reference(buyers, n_sellers, c["T"],
                                        seller_stage_handler=wl.seller_handler_avg_price)
            tr = harness.run_generic(env, harness.EpisodeClock([coords]), actions[e], mask[e],
                                     wl.OBS_DIM, state_fn=wl.state, convert=wl.to_action(env))
        tr["messages"] = []
        per_env.append(tr)
    out = pack_generic(per_env, actions, mask, c["seed"], 0, {})
    out["buyers"] = np.array(buyers, np.float64)
    out["n_sellers"] = np.int64(n_sellers)
    np.savez_compressed(os.path.join(GOLDEN, "simple_market_handler_reference.npz"), **out)
    print("simple_market_handler_reference.npz", {k: v.shape for k, v in out.items()})


FSM_HANDLER_FUZZ_CASES = 40
FSM_WIDE_FUZZ_CASES = 16
FSM_FLOAT_FUZZ_CASES = 32
FSM_WAITING_FUZZ_CASES = 40
MOCK_ENV_FUZZ_CASES = 60
FSM_ORDER_FUZZ_SEEDS = (1103, 1308, 1334, 1732, 1734, 2051, 2484, 2509, 2614, 2937, 2948, 3054, 3380)


def gen_fsm_handler_fuzz_reference() -> None:
    """fsm_handler_fuzz_reference.json: tests/kat_scenarios.py:run_random_handler_fsm (random
    FiniteStateMachineEnvs with env stage handlers, fsm.py:294-307) executed by the UNMODIFIED
    reference for case seeds 0..39: per step the stage, observations, rewards, done flags, the
    mock agents' call counters and the echo agents' message counters; or the exception type."""
    import json

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from tests import kat_scenarios as kats

    from .workloads import mock

    K = mock.build_classes(ref_shim.import_reference())
    out = {str(s): kats.run_random_handler_fsm(K, s) for s in range(FSM_HANDLER_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "fsm_handler_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("fsm_handler_fuzz_reference.json", len(out), "cases; raising:", raised)
    # the same with compound handlers (if / elif / else chains of 1-2 comparisons per branch)
    out = {str(s): kats.run_random_handler_fsm(K, s, compound=True)
           for s in range(FSM_HANDLER_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "fsm_compound_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("fsm_compound_fuzz_reference.json", len(out), "cases; raising:", raised)
    # ... with float32 comparisons in the chains (an echo agent's float32 `level`)
    out = {str(s): kats.run_random_handler_fsm(K, s, floats=True) for s in range(FSM_FLOAT_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "fsm_float_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("fsm_float_fuzz_reference.json", len(out), "cases; raising:", raised)
    # ... the float32 cases of a 2 500-seed device-vs-oracle campaign (tools/fuzz_campaign.py) whose
    # outcome depends on the ORDER in which a stage's agents act: the reference walks
    # FSMStage.acting_agents in list order (fsm.py:276-277), not in network order, and the echo
    # agents' float32 `level` recurrence is sensitive to the order of a receiver's batch
    out = {str(s): kats.run_random_handler_fsm(K, s, floats=True) for s in FSM_ORDER_FUZZ_SEEDS}
    with open(os.path.join(GOLDEN, "fsm_order_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    print("fsm_order_fuzz_reference.json", len(out), "cases")
    # ... handlers that do NOT resolve although mail was sent: the mail waits in the resolver for a
    # later step's resolve_network() (fsm.py:280-283)
    out = {str(s): kats.run_random_handler_fsm(K, s, waiting=True) for s in range(FSM_WAITING_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "fsm_waiting_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("fsm_waiting_fuzz_reference.json", len(out), "cases; raising:", raised)
    # random PhantomEnv / StackelbergEnv env classes (message traces, round limits, bad edges,
    # shuffled batches under the contract shuffle): tests/kat_scenarios.py:run_mock_env
    out = {str(s): kats.run_mock_env(K, s) for s in range(MOCK_ENV_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "mock_env_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("mock_env_fuzz_reference.json", len(out), "cases; raising:", raised,
          "shuffled:", sum(bool(t[0][1]) for t in out.values()))
    # ... and on env classes wider than a warp (33..120 agents): the block engine's fixture
    out = {str(s): kats.run_random_handler_fsm(K, s, wide=True) for s in range(FSM_WIDE_FUZZ_CASES)}
    with open(os.path.join(GOLDEN, "fsm_wide_fuzz_reference.json"), "w") as f:
        json.dump(out, f, separators=(",", ":"))
    raised = sorted(int(s) for s, t in out.items() if t[-1][0] == "raise")
    print("fsm_wide_fuzz_reference.json", len(out), "cases; raising:", raised,
          "bytes:", os.path.getsize(os.path.join(GOLDEN, "fsm_wide_fuzz_reference.json")))


def gen_digital_ads_reference(only=slice(None)) -> None:
    """The reference's examples/environments/digital_ads_market/digital_ads_market.py, UNMODIFIED
    (oracle/workloads/digital_ads.py:build_reference), under the contract RNG:
      digital_ads_reference.npz       2 + 2 + 2 advertisers, first-price auction (8 agents)
      digital_ads_wide_reference.npz  10 + 10 + 10 advertisers, second-price auction (32 agents)
      digital_ads_full_reference.npz  40 + 40 + 40 advertisers, the example's SHIPPED size
                                      (digital_ads_market.py:687-689; 122 agents), first price"""
    from .workloads import digital_ads as wl

    for name, per_theme, strategy, n_env, n_ep, seed in (
            ("digital_ads_reference.npz", 2, "first", 8, 2, 20261026),
            ("digital_ads_wide_reference.npz", 10, "second", 3, 2, 20261027),
            ("digital_ads_full_reference.npz", 40, "first", 2, 2, 20261028))[only]:
        theme = {"travel": per_theme, "tech": per_theme, "sport": per_theme}
        budgets = ([(5.0, 15.001, 5.0, 15.0)] * per_theme + [(7.0, 17.001, 7.0, 17.0)] * per_theme +
                   [(10.0, 20.001, 10.0, 20.0)] * per_theme)
        S, T = 3 * per_theme, 20
        actions, mask = wl.actions_for(n_env, n_ep, T, S, seed % 1000)
        per_env = []
        for e in range(n_env):
            coords = wl.Coords(seed, e)
            with wl.contract_rng(coords):
                env = wl.build_reference(theme, budgets, T, strategy)
                tr = harness.run_generic(env, harness.EpisodeClock([coords]), actions[e], mask[e],
                                         wl.OBS_DIM, state_fn=wl.state, flatten=wl.flatten_obs)
            tr["messages"] = []
            per_env.append(tr)
        out = pack_generic(per_env, actions, mask, seed, 0, {})
        out["per_theme"], out["budgets"] = np.int64(per_theme), np.array(budgets, np.float64)
        out["second_price"] = np.int64(strategy == "second")
        np.savez_compressed(os.path.join(GOLDEN, name), **out)
        print(name, {k: v.shape for k, v in out.items()}, "terminated:",
              int((out["term"] == 1).sum()), "clicks:", float(out["reward"].sum()))


def main() -> int:
    if not ref_shim.reference_available():
        print("reference not available")
        return 2
    os.makedirs(GOLDEN, exist_ok=True)
    gen_supply_chain_reference()
    gen_market_reference()
    gen_stackelberg_reference()
    gen_dense_reference()
    gen_supply_chain2_reference()
    gen_shuffle_and_stochastic_reference()
    gen_simple_market_reference()
    gen_simple_market_handler_reference()
    gen_fsm_handler_fuzz_reference()
    gen_digital_ads_reference()
    return 0


if __name__ == "__main__":
    sys.exit(main())
