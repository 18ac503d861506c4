"""Oracle vs fixtures generated from the UNMODIFIED reference (oracle/make_golden.py).

CPU only.  These are what make the oracle trustworthy as the checker of the CUDA path."""
import json
import os

import numpy as np
import pytest

import oracle.phantom_oracle as po
from oracle import harness, rng, vectorised
from oracle.workloads import supply_chain as wl

STEP_KEYS = ["obs", "reward", "term", "trunc", "all_term", "all_trunc", "state"]


@pytest.fixture(scope="module")
def sc_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "supply_chain_reference.npz"))


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    for ctr, key, want in rng.KAT:
        assert rng.philox4x32(ctr, key) == want
        got = rng.philox4x32_np(*[np.array([c]) for c in ctr], *[np.array([k]) for k in key])
        assert tuple(int(x[0]) for x in got) == want


def test_contract_scalar_equals_vectorised():
    env = np.arange(50)
    for idx in (0, 3, 4, 5, 9, 14):
        v = rng.d24_np(77, env, 2, 13, 1, idx)
        s = [rng.d24(77, int(e), 2, 13, 1, idx) for e in env]
        assert v.tolist() == s and max(s) < 2**24
    assert all(0 <= rng.randint(5, d) < 5 for d in (0, 1, 2**23, 2**24 - 1))
    # the five draws of a block: four high 24-bit fields + the three low bytes of w0..w2
    w = rng.philox4x32((3, 0, 1, 0), (9, 0))
    assert [rng.d24(9, 3, 0, 1, 0, k) for k in range(5)] == [
        w[0] >> 8, w[1] >> 8, w[2] >> 8, w[3] >> 8,
        ((w[0] & 255) << 16) | ((w[1] & 255) << 8) | (w[2] & 255)]
    assert rng.d24(9, 3, 0, 1, 0, 5) == rng.philox4x32((3, 0, 1, 1), (9, 0))[0] >> 8


def test_packed_draws():
    """Contract v3: base-n digits of Philox words; scalar == vectorised == PackedStream, one
    block serves four steps when K <= kpw, and the digits are uniform and uncorrelated."""
    assert [rng.digits_per_word(n) for n in (2, 5, 16, 255, 256)] == [16, 6, 4, 2, 2]
    for n, K in [(5, 5), (5, 7), (255, 30), (2, 20), (16, 9)]:
        a = rng.packed_randint_np(7, np.arange(3)[:, None], 1, np.arange(9)[None, :], 0, n, K)
        assert a.shape == (3, 9, K) and a.min() >= 0 and a.max() < n
        for e in range(3):
            st = rng.PackedStream(7, e, 0, n, K)
            for step in range(9):
                st.begin(1, step)
                want = [rng.packed_randint(7, e, 1, step, 0, n, K, i) for i in range(K)]
                assert a[e, step].tolist() == want == [st.randint(n) for _ in range(K)]
    # word layout: K = n = 5 -> draw i of step s = digit i of word s; word s = w[s & 3] of
    # block s >> 2 with the 0x8000 marker in the counter's stream word
    w = rng.philox4x32((3, 2, 6 >> 2, 0x8000), (9, 0))
    x, digits = w[6 & 3], []
    for _ in range(5):
        x *= 5
        digits.append(x >> 32)
        x &= 0xFFFFFFFF
    assert [rng.packed_randint(9, 3, 2, 6, 0, 5, 5, i) for i in range(5)] == digits
    big = rng.packed_randint_np(3, np.arange(20000)[:, None], 0, np.arange(1, 11)[None, :], 0, 5, 5)
    freq = np.bincount(big.ravel(), minlength=5) / big.size
    assert np.abs(freq - 0.2).max() < 0.003
    flat = big.reshape(-1, 5).astype(np.float64)
    c = np.corrcoef(flat.T)
    assert np.abs(c - np.eye(5)).max() < 0.01


def test_survey_probe_trace():
    """SURVEY.md 3.6 / 8c probe: np.random.seed(0) orders (4,0,3,3,3 then ...) are not
    reproducible under the contract stream, but the dynamics are: feed the same orders."""

    class Fixed:
        def __init__(self, seq):
            self.seq = list(seq)

        def randint(self, n):
            return self.seq.pop(0)

    # orders drawn by the reference with np.random.seed(0): 4,0,3,3,3 | 1,3,2,4,0 | 0,4,2,1,0
    st = Fixed([4, 0, 3, 3, 3, 1, 3, 2, 4, 0, 0, 4, 2, 1, 0])
    env = wl.build(po, st)
    env.reset()
    s = env.step({"SHOP": np.array([69.64692], np.float32)})
    assert s.observations["SHOP"].tolist() == np.array([0.70, 0.0, 0.52], np.float32).tolist()
    assert s.rewards["SHOP"] == -7.0
    s = env.step({"SHOP": np.array([28.613934], np.float32)})
    assert s.observations["SHOP"].tolist() == np.array([0.89, 0.40, 0.0], np.float32).tolist()
    assert s.rewards["SHOP"] == 1.0999999999999996
    s = env.step({"SHOP": np.array([22.685144], np.float32)})
    assert s.observations["SHOP"].tolist() == np.array([0.93, 0.28, 0.0], np.float32).tolist()
    assert s.rewards["SHOP"] == -2.3000000000000007


def test_object_oracle_matches_reference_golden(sc_golden):
    g = sc_golden
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    for e in range(A.shape[0]):
        st = wl.order_stream(seed, e)
        env = wl.build(po, st, enable_tracking=e < 3)
        tr = harness.run_supply_chain(env, harness.EpisodeClock([st]), A[e], M[e], track=e < 3)
        for k in ["reset_obs"] + STEP_KEYS:
            assert np.array_equal(tr[k], g[k][e]), (e, k)
        if e < 3:  # exact global message order == Resolver.tracked_messages of the reference
            gm = g["messages"]
            assert np.array_equal(tr["messages"], gm[gm[:, 0] == e][:, 1:])


def test_vectorised_oracle_matches_reference_golden(sc_golden):
    g = sc_golden
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    v = vectorised.SupplyChainVec(A.shape[0], seed)
    for ep in range(A.shape[1]):
        assert np.array_equal(v.reset(), g["reset_obs"][:, ep])
        for t in range(A.shape[2]):
            s = v.step(A[:, ep, t], M[:, ep, t])
            for k in STEP_KEYS:
                assert np.array_equal(s[k], g[k][:, ep, t]), (ep, t, k)


def test_golden_covers_quirks(sc_golden):
    g = sc_golden
    assert (g["state"][..., 0] < 0).any(), "negative stock (no lower clamp) not exercised"
    assert (g["action_mask"] == 0).any(), "missing-action fallback not exercised"
    a = g["actions"][..., 0]
    assert ((a - np.floor(a)) == 0.5).any(), "round-half-even not exercised"
    assert (g["reset_obs"][:, 1, 1:] != 0).any(), "sales/missed carry-over across reset not exercised"
    assert g["all_trunc"][:, :, -1].all() and not g["all_trunc"][:, :, :-1].any()


def test_pinned_record(golden_dir):
    """oracle/pin_against_reference.py ran the reference's own hot-path tests against the
    reference (through the shim) and against the oracle; both must have been green, and
    the oracle must not have changed since."""
    from oracle import pin_against_reference as pin

    rec = json.load(open(os.path.join(golden_dir, "PINNED.json")))
    assert rec["reference"]["returncode"] == 0 and rec["oracle"]["returncode"] == 0
    assert rec["oracle"]["counts"]["passed"] == rec["reference"]["counts"]["passed"] >= 64
    assert rec["oracle_sha256"] == pin.oracle_fingerprint(), (
        "oracle/phantom_oracle changed: re-run `python -m oracle.pin_against_reference`")


def test_object_oracle_matches_market_golden(golden_dir):
    """C3 workload: oracle restatement == the reference running the same env definition."""
    from oracle import make_golden
    from oracle.workloads import market

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "market_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    for e in range(2):
        st = rng.StepStream(seed, e, market.STREAM_TAKER_VALUE)
        env = market.build(po, st, enable_tracking=e < 1)
        tr = harness.run_generic(env, harness.EpisodeClock([st]), A[e], M[e], 3, track=e < 1,
                                 state_fn=make_golden.market_state)
        assert_oracle_trace_equal(tr, g, e)
        if e < 1:
            rows = [(0, ep, t, s, r, make_golden.MESSAGE_TYPE_IDS[n], v0, v1)
                    for (ep, t, s, r, n, v0, v1) in tr["messages"]]
            assert np.array_equal(np.asarray(rows, np.int64), g["messages"])


def test_object_oracle_matches_stackelberg_golden(golden_dir):
    """C4 workload: oracle restatement == the reference running the same env definition."""
    from oracle.workloads import stackelberg as wl

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "stackelberg_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    rows = []
    for e in range(6):
        st = rng.StepStream(seed, e, wl.STREAM_FOLLOWER_VALUE)
        env = wl.build(po, st, enable_tracking=e < 4)
        tr = harness.run_generic(env, harness.EpisodeClock([st]), A[e], M[e], 2, track=e < 4,
                                 state_fn=wl.state)
        assert_oracle_trace_equal(tr, g, e)
        if e < 4:
            rows += [(e, ep, t, s, r, wl.MESSAGE_TYPE_IDS[n], v0, v1)
                     for (ep, t, s, r, n, v0, v1) in tr["messages"]]
    assert np.array_equal(np.asarray(rows, np.int64), g["messages"])


def test_object_oracle_matches_dense_goldens(golden_dir):
    """C5 workload: oracle restatement == the reference running the same env definition (the
    sparse 12-agent fixture incl. message order; one env of the 128-agent fixture)."""
    from oracle.workloads import dense as wl

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "dense12_reference.npz"))
    A, M = g["actions"], g["action_mask"]
    rows = []
    for e in range(3):
        env = wl.build(po, n_agents=12, adjacency=g["adjacency"], enable_tracking=True)
        tr = harness.run_generic(env, harness.EpisodeClock([]), A[e], M[e], 3, track=True,
                                 state_fn=wl.state)
        assert_oracle_trace_equal(tr, g, e)
        rows += [(e, ep, t, s, r, wl.MESSAGE_TYPE_IDS[n], v0, v1)
                 for (ep, t, s, r, n, v0, v1) in tr["messages"]]
    gm = g["messages"]
    assert np.array_equal(np.asarray(rows, np.int64), gm[gm[:, 0] < 3])
    g = np.load(os.path.join(golden_dir, "dense128_reference.npz"))
    env = wl.build(po, n_agents=128)
    tr = harness.run_generic(env, harness.EpisodeClock([]), g["actions"][0, :1], g["action_mask"][0, :1],
                             3, state_fn=wl.state)
    for k in ["obs", "reward", "state", "obs_mask", "reward_mask", "all_done"]:
        assert np.array_equal(tr[k][0], g[k][0, 0]), k


def test_object_oracle_matches_supply_chain2_golden(golden_dir):
    """Tutorial-2 env with supertypes: oracle restatement (incl. its UniformFloatSampler) == the
    reference, both drawing np.random.uniform through the contract stream."""
    from oracle.workloads import supply_chain2 as wl

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "supply_chain2_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    for e in range(4):
        streams = {s: rng.StepStream(seed, e, s)
                   for s in (wl.STREAM_ORDER, wl.STREAM_SAMPLER, wl.STREAM_SHOP_CHOICE)}
        ws = []
        with harness.patched_np_uniform(streams[wl.STREAM_SAMPLER]):
            env = wl.build(po, streams, po.utils.samplers.UniformFloatSampler, enable_tracking=e < 3)
            tr = harness.run_generic(
                env, harness.EpisodeClock(list(streams.values())), A[e], M[e], 4, track=e < 3,
                state_fn=lambda env: (ws.append(wl.weights(env)), wl.state(env))[1])
        assert_oracle_trace_equal(tr, g, e)
        assert np.array_equal(np.array(ws).reshape(g["weights"][e].shape), g["weights"][e])


def test_object_oracle_matches_market_shuffle_golden(golden_dir):
    """BatchResolver(shuffle_batches=True): oracle restatement == reference under the contract's
    Fisher-Yates, and the shuffle is observable (differs from the unshuffled run)."""
    from oracle import make_golden
    from oracle.workloads import market

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "market_shuffle_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    e = 0
    st = rng.StepStream(seed, e, market.STREAM_TAKER_VALUE)
    env = market.build(po, st, enable_tracking=True, shuffle_batches=True)
    clock = harness.EpisodeClock([st])
    slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
    with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
        tr = harness.run_generic(env, clock, A[e], M[e], 3, track=True,
                                 state_fn=make_golden.market_state)
    assert_oracle_trace_equal(tr, g, e)
    rows = [(0, ep, t, s, r, make_golden.MESSAGE_TYPE_IDS[n], v0, v1)
            for (ep, t, s, r, n, v0, v1) in tr["messages"]]
    assert np.array_equal(np.asarray(rows, np.int64), g["messages"])
    # same env without shuffling ends up in a different state
    st = rng.StepStream(seed, e, market.STREAM_TAKER_VALUE)
    plain = harness.run_generic(market.build(po, st), harness.EpisodeClock([st]), A[e][:1], M[e][:1],
                                3, state_fn=make_golden.market_state)
    assert not np.array_equal(plain["state"][0], g["state"][e, 0])


def test_object_oracle_matches_stochastic_network_golden(golden_dir):
    """StochasticNetwork (resampled per episode, ignore_connection_errors) + shuffled batches:
    oracle restatement == reference, graphs included; edge frequencies follow the rates."""
    from oracle.workloads import supply_chain2 as wl

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "supply_chain2_stochastic_reference.npz"))
    seed, A, M, rates = int(g["seed"]), g["actions"], g["action_mask"], tuple(g["rates"])
    T = A.shape[2]
    for e in range(5):
        streams = {s: rng.StepStream(seed, e, s)
                   for s in (wl.STREAM_ORDER, wl.STREAM_SAMPLER, wl.STREAM_SHOP_CHOICE,
                             wl.STREAM_CONNECTIVITY)}
        adj = []
        with harness.patched_np_uniform(streams[wl.STREAM_SAMPLER]), \
                harness.patched_np_random(streams[wl.STREAM_CONNECTIVITY]):
            env = wl.build(po, streams, po.utils.samplers.UniformFloatSampler, num_steps=T,
                           enable_tracking=e < 4, rates=rates, shuffle_batches=True)
            clock = harness.EpisodeClock(list(streams.values()))
            slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
            with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
                tr = harness.run_generic(
                    env, clock, A[e], M[e], 4, track=e < 4,
                    state_fn=lambda env: (adj.append(wl.adjacency(env)), wl.state(env))[1])
        assert_oracle_trace_equal(tr, g, e)
        assert np.array_equal(np.array(adj).reshape(-1, T, 8, 8)[:, 0], g["adjacency"][e])
    # the contract draw reproduces the graphs directly: connection c exists iff d24 * 2^-24 < rate
    base = [(1, 0), (2, 0)] + [(s, c) for s in (1, 2) for c in range(3, 8)]
    rate = [rates[0]] * 2 + [rates[1]] * 10
    for e in range(g["adjacency"].shape[0]):
        for ep in range(g["adjacency"].shape[1]):
            want = np.zeros((8, 8), np.uint8)
            for c, (u, v) in enumerate(base):
                if rng.d24(seed, e, ep, 0, wl.STREAM_CONNECTIVITY, c) / 16777216.0 < rate[c]:
                    want[u, v] = want[v, u] = 1
            assert np.array_equal(want, g["adjacency"][e, ep]), (e, ep)


@pytest.mark.parametrize("name", ["simple_market_reference.npz", "simple_market_wide_reference.npz",
                                  "simple_market_9s_reference.npz",
                                  "simple_market_block_reference.npz"])
def test_object_oracle_matches_simple_market_golden(golden_dir, name):
    """The reference's own simple_market example (env-level post_message_resolution + custom
    EnvView field, two-stage FSM, three RNG call sites): oracle restatement == the unmodified
    example modules run by the reference, float64 state included."""
    from oracle.workloads import simple_market as sm

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, name))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    buyers, n_sellers, T = [tuple(b) for b in g["buyers"]], int(g["n_sellers"]), A.shape[2]
    for e in range(A.shape[0]):
        coords = sm.Coords(seed, e)
        with sm.contract_rng(coords, {f"b{i + 1}": i for i in range(len(buyers))}):
            env, _ = sm.build(po, po.utils.samplers.UniformFloatSampler, buyers, n_sellers, T)
            env.network.resolver.enable_tracking = e < 2
            tr = harness.run_generic(env, harness.EpisodeClock([coords]), A[e], M[e], sm.OBS_DIM,
                                     state_fn=sm.state, convert=sm.to_action(env),
                                     track="raw" if e < 2 else False)
        assert_oracle_trace_equal(tr, g, e)
        if e < 2:  # Resolver.tracked_messages: global order, float64 prices
            slot_rows = [(ep, t, s, r, {"Price": 0, "Order": 1}[name], v0)
                         for (ep, t, s, r, name, v0, v1) in tr["messages"]]
            gm = g["messages"]
            assert np.array_equal(np.asarray(slot_rows, np.float64), gm[gm[:, 0] == e][:, 1:])
    # the fixture exercises what it is meant to: ties between sellers, withheld buyer actions,
    # None rewards at the first buyer observation, and a moving avg_price in the sellers' obs
    assert (g["reward_mask"] == 2).any() and (M == 0).any()
    assert len(np.unique(g["state"][..., -1, 0])) > 10


def test_object_oracle_matches_simple_market_handler_golden(golden_dir):
    """Handler-driven FSM transitions (fsm.py:294-307): the oracle restatement of simple_market
    with a Python env handler on the Sellers stage == the UNMODIFIED example classes run by the
    reference with the same handler (oracle/make_golden.py:gen_simple_market_handler_reference).
    The GPU test test_simple_market_handler_driven_on_env_word compares the device (StageRule on
    the env-level avg_price word) with this oracle on the same cast and action tape."""
    from oracle.workloads import simple_market as sm

    from .generic_parity import assert_oracle_trace_equal

    g = np.load(os.path.join(golden_dir, "simple_market_handler_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    buyers, n_sellers, T = [tuple(b) for b in g["buyers"]], int(g["n_sellers"]), A.shape[2]
    A2, M2 = sm.actions_for(A.shape[0], A.shape[1], T, len(buyers), n_sellers, 9)
    assert seed == 5 and np.array_equal(A, A2) and np.array_equal(M, M2)  # the GPU test's tape
    for e in range(A.shape[0]):
        coords = sm.Coords(seed, e)
        with sm.contract_rng(coords, {f"b{i + 1}": i for i in range(len(buyers))}):
            env, _ = sm.build(po, po.utils.samplers.UniformFloatSampler, buyers, n_sellers, T,
                              seller_stage_handler=sm.seller_handler_avg_price)
            tr = harness.run_generic(env, harness.EpisodeClock([coords]), A[e], M[e], sm.OBS_DIM,
                                     state_fn=sm.state, convert=sm.to_action(env))
        assert_oracle_trace_equal(tr, g, e)
    # both outcomes of the handler occur: steps after which the sellers observe again (the FSM
    # stayed in / returned to Sellers) and steps after which only buyers do
    sellers_next = g["obs_mask"][..., len(buyers):].any(axis=-1)
    assert sellers_next.any() and (~sellers_next).any()
    # and the handler made envs leave the example's strict Sellers/Buyers alternation
    assert (sellers_next[:, :, :-1] & sellers_next[:, :, 1:]).any()


@pytest.mark.parametrize("name", ["digital_ads_reference.npz", "digital_ads_wide_reference.npz",
                                  "digital_ads_full_reference.npz"])
def test_oracle_port_runs_the_digital_ads_example(golden_dir, name):
    """The UNMODIFIED example file executed on the oracle port (its `import phantom` bound to
    oracle.phantom_oracle) reproduces the fixture the same file produced on the reference: pins
    the port's StochasticNetwork, FSM caches, handle_batch override dispatch, done-agent dropout
    and env-managed clipped samplers.  Needs the example source, i.e. the build container."""
    import importlib.util
    import sys

    from oracle import ref_shim
    from oracle.workloads import digital_ads as wl

    from .generic_parity import assert_oracle_trace_equal

    path = os.path.join(ref_shim.REFERENCE_ROOT,
                        "examples/environments/digital_ads_market/digital_ads_market.py")
    if not os.path.exists(path):
        pytest.skip("reference examples are only available in the build container")
    ref_shim.install(with_reference=False)  # gymnasium & co stubs only
    import types
    import typing

    # the file also defines metric classes at module level (:594-684, not on the step path): give
    # the port's namespace the two names they derive from
    port = types.ModuleType("phantom")
    port.__dict__.update(po.__dict__)
    T = typing.TypeVar("T")
    class Metric(typing.Generic[T]):
        pass

    port.metrics = types.SimpleNamespace(Metric=Metric, SimpleAgentMetric=lambda *a, **k: None)
    saved = {k: sys.modules.get(k) for k in ("phantom",)}
    sys.modules["phantom"] = port
    try:
        spec = importlib.util.spec_from_file_location("_port_digital_ads", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    g = np.load(os.path.join(golden_dir, name))
    seed, A, M, per_theme = int(g["seed"]), g["actions"], g["action_mask"], int(g["per_theme"])
    theme = {"travel": per_theme, "tech": per_theme, "sport": per_theme}
    for e in range(min(A.shape[0], 4)):
        coords = wl.Coords(seed, e)
        with wl.contract_rng(coords):
            st = {f"ADV_{i + 1}": mod.AdvertiserAgent.Supertype(
                budget=po.utils.samplers.UniformFloatSampler(*b)) for i, b in enumerate(g["budgets"])}
            env = mod.DigitalAdsEnv(num_steps=A.shape[2], num_agents_theme=theme, agent_supertypes=st)
            env.agents["ADX"].strategy = "second" if int(g["second_price"]) else "first"
            tr = harness.run_generic(env, harness.EpisodeClock([coords]), A[e], M[e], wl.OBS_DIM,
                                     state_fn=wl.state, flatten=wl.flatten_obs)
        assert_oracle_trace_equal(tr, g, e)
