"""Parity of the CUDA supply-chain path (through the C ABI) against the oracle and the
fixtures generated from the unmodified reference.  Runs on the B200 box."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

from oracle import harness, rng, vectorised  # noqa: E402
import oracle.phantom_oracle as po  # noqa: E402
from oracle.workloads import supply_chain as wl  # noqa: E402

STEP_KEYS = ["obs", "reward", "term", "trunc", "all_term", "all_trunc", "state"]


@pytest.fixture(scope="module")
def ph():
    import phantom_b200

    return phantom_b200


@pytest.fixture(scope="module")
def sc():
    from phantom_b200.envs import supply_chain

    return supply_chain


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "supply_chain_reference.npz"))


def shop_state(env):
    shop = env.agents["SHOP"]
    return np.stack([np.atleast_1d(env.agent_column(shop, w)) for w in range(4)], axis=1)


def assert_step_equal(out, ref, ctx=""):
    """out: BatchStep (device) for one step; ref: dict from an oracle.  Bit-exact."""
    obs = out.observations.cpu().numpy()[:, 0, :]
    assert np.array_equal(obs, ref["obs"]), f"obs {ctx}"
    # reward: float32(reference float64) -- tolerance of the spec is 1e-5 rel; we are exact
    rew = out.rewards.cpu().numpy()[:, 0]
    assert np.array_equal(rew, ref["reward"].astype(np.float32)), f"reward {ctx}"
    assert np.array_equal(out.terminations.cpu().numpy()[:, 0], ref["term"]), f"term {ctx}"
    assert np.array_equal(out.truncations.cpu().numpy()[:, 0], ref["trunc"]), f"trunc {ctx}"
    ad = out.all_done.cpu().numpy()
    assert np.array_equal(ad[:, 0], ref["all_term"]) and np.array_equal(ad[:, 1], ref["all_trunc"]), ctx
    assert out.obs_mask.cpu().numpy().all() and (out.reward_mask.cpu().numpy() == 1).all()


def test_native_library_is_loaded(ph):
    from phantom_b200 import _lib

    assert _lib.lib.phx_device_count() >= 1
    assert os.path.basename(_lib.LIB_PATH) == "libphx.so"


@pytest.mark.parametrize("exec_mode", ["fast", "thread", "queue"])
def test_single_step_api_matches_reference_golden(sc, golden, exec_mode):
    g = golden
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    n_env, n_ep, T = A.shape[:3]
    env = sc.SupplyChainEnv(num_envs=n_env, seed=seed, exec_mode=exec_mode)
    for ep in range(n_ep):
        obs, mask = env.reset_batch()
        assert np.array_equal(obs.cpu().numpy()[:, 0, :], g["reset_obs"][:, ep])
        assert mask.cpu().numpy().all()
        for t in range(T):
            out = env.step_batch(A[:, ep, t].reshape(n_env, 1, 1), M[:, ep, t].reshape(n_env, 1))
            ref = {k: g[k][:, ep, t] for k in STEP_KEYS}
            assert_step_equal(out, ref, f"ep {ep} t {t}")
            assert np.array_equal(shop_state(env), ref["state"]), (ep, t)
    env.check_errors()
    env.close()


@pytest.mark.parametrize("exec_mode", ["fast", "thread", "queue"])
def test_message_trace_matches_reference_tracked_messages(sc, golden, exec_mode):
    """Bit-exact routing: the device trace equals Resolver.tracked_messages of the reference
    (global push order over both rounds) for every step of 3 envs x 2 episodes."""
    g = golden
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    gm = g["messages"]  # (env, ep, t, sender, recv, type, v0, v1)
    n_env = 3
    env = sc.SupplyChainEnv(num_envs=n_env, seed=seed, enable_tracking=True, exec_mode=exec_mode)
    for ep in range(A.shape[1]):
        env.reset_batch()
        for t in range(A.shape[2]):
            env.step_batch(A[:n_env, ep, t].reshape(n_env, 1, 1), M[:n_env, ep, t].reshape(n_env, 1))
            counts, rows = env.tracked_messages_batch()
            for e in range(n_env):
                want = gm[(gm[:, 0] == e) & (gm[:, 1] == ep) & (gm[:, 2] == t)][:, 3:7].astype(np.int64)
                r = rows[e, : counts[e]]
                got = np.stack([r[:, 0] & 0xFF, (r[:, 0] >> 8) & 0xFF, (r[:, 0] >> 16) & 0xFF, r[:, 1]], 1)
                assert np.array_equal(got, want), (e, ep, t)
    env.close()


@pytest.mark.parametrize("exec_mode", ["fast", "thread", "queue"])
def test_rollout_equals_single_steps(sc, exec_mode):
    E, T, seed = 1000, 37, 5
    r = np.random.RandomState(0)
    A = r.uniform(-20, 150, size=(T, E, 1, 1)).astype(np.float32)
    a = sc.SupplyChainEnv(num_envs=E, seed=seed, exec_mode=exec_mode)
    b = sc.SupplyChainEnv(num_envs=E, seed=seed, exec_mode=exec_mode)
    a.reset_batch(); b.reset_batch()
    ro = b.rollout_batch(A)
    for t in range(T):
        out = a.step_batch(A[t])
        for x, y in zip(out, ro):
            assert torch.equal(x, y[t]), t
    assert np.array_equal(shop_state(a), shop_state(b))
    assert np.array_equal(a.field(0, np.int32), b.field(0, np.int32))
    a.close(); b.close()


def test_full_size_against_vectorised_oracle(sc):
    """BASELINE config C2: 65 536 envs x one full 100-step episode, bit-exact against the
    numpy restatement (which is itself pinned to the reference fixtures)."""
    E, T, seed = 65536, 100, 20261017
    r = np.random.RandomState(1)
    A = r.uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
    A[r.randint(0, T, 500), r.randint(0, E, 500)] *= -1.0
    env = sc.SupplyChainEnv(num_envs=E, seed=seed)
    v = vectorised.SupplyChainVec(E, seed)
    obs0, _ = env.reset_batch()
    assert np.array_equal(obs0.cpu().numpy()[:, 0], v.reset())
    ro = env.rollout_batch(A)
    obs = ro.observations.cpu().numpy()[:, :, 0]
    rew = ro.rewards.cpu().numpy()[:, :, 0]
    ad = ro.all_done.cpu().numpy()
    for t in range(T):
        ref = v.step(A[t, :, 0, 0])
        assert np.array_equal(obs[t], ref["obs"]), t
        assert np.array_equal(rew[t], ref["reward"].astype(np.float32)), t
        assert np.array_equal(ad[t, :, 1], ref["all_trunc"]), t
    assert np.array_equal(shop_state(env), ref["state"])
    env.check_errors()
    env.close()


@pytest.mark.parametrize("E", [1, 31, 32, 33, 96, 160, 224])
def test_small_and_ragged_env_counts_against_vectorised_oracle(sc, E):
    """Env counts around the block / warp sizes of the fast kernel (E % 64 == 32 once selected
    the vector action copy for a block whose second warp holds no env: ADVICE r1)."""
    T, seed = 23, 77
    A = np.random.RandomState(E).uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
    env = sc.SupplyChainEnv(num_envs=E, seed=seed)
    assert env.exec_name.startswith("fast")
    v = vectorised.SupplyChainVec(E, seed)
    obs0, _ = env.reset_batch()
    assert np.array_equal(obs0.cpu().numpy()[:, 0], v.reset())
    ro = env.rollout_batch(A)
    obs = ro.observations.cpu().numpy()[:, :, 0]
    rew = ro.rewards.cpu().numpy()[:, :, 0]
    for t in range(T):
        ref = v.step(A[t, :, 0, 0])
        assert np.array_equal(obs[t], ref["obs"]), t
        assert np.array_equal(rew[t], ref["reward"].astype(np.float32)), t
    assert np.array_equal(shop_state(env), ref["state"])
    env.check_errors()
    env.close()


@pytest.mark.parametrize("num_steps,T,E", [(10, 57, 1000), (100, 100, 4096), (7, 40, 333), (4, 19, 64),
                                           (3, 11, 100), (24, 131, 2048), (100, 250, 1024)])
def test_fast_kernel_variants_agree(sc, monkeypatch, num_steps, T, E):
    """The three schedule-specialised kernels (PHX_SC_KERNEL = 1: round-1 thread-per-env, 2:
    closed-form fill + aligned groups, 3: time-parallel, with 4 or 2 warps per 32 envs, 4: two
    lanes per env -- used when the env count is a multiple of 64, else 2) produce
    identical planes and state, with auto-reset wraps inside the launch, launches that start at
    every clock phase (three consecutive rollouts of T steps) and negative actions."""
    r = np.random.RandomState(T)
    A = [r.uniform(-30, 130, size=(T, E, 1, 1)).astype(np.float32) for _ in range(3)]
    results = []
    for variant, warps in ((1, 4), (2, 4), (3, 4), (3, 2), (4, 4)):
        monkeypatch.setenv("PHX_SC_KERNEL", str(variant))
        monkeypatch.setenv("PHX_SC_WARPS", str(warps))
        env = sc.SupplyChainEnv(num_envs=E, seed=12, num_steps=num_steps, auto_reset=True)
        env.reset_batch()
        outs = [[x.clone() for x in env.rollout_batch(a)] for a in A]
        results.append((outs, shop_state(env), env.field(0, np.int32), env.field(1, np.int32)))
        env.check_errors()
        env.close()
    base = results[0]
    for other in results[1:]:
        for oa, ob in zip(base[0], other[0]):
            for x, y in zip(oa, ob):
                assert torch.equal(x, y)
        assert np.array_equal(base[1], other[1])
        assert np.array_equal(base[2], other[2]) and np.array_equal(base[3], other[3])


def test_sharding_invariance(sc):
    """Results do not depend on how envs are split over handles (multi-GPU sharding uses
    env_offset; here two handles on one GPU)."""
    E, T, seed = 512, 20, 9
    A = np.random.RandomState(2).uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
    whole = sc.SupplyChainEnv(num_envs=E, seed=seed)
    lo = sc.SupplyChainEnv(num_envs=E // 2, seed=seed, env_offset=0)
    hi = sc.SupplyChainEnv(num_envs=E // 2, seed=seed, env_offset=E // 2)
    for e in (whole, lo, hi):
        e.reset_batch()
    w = whole.rollout_batch(A)
    l = lo.rollout_batch(A[:, : E // 2])
    h = hi.rollout_batch(A[:, E // 2:])
    for x, y, z in zip(w, l, h):
        assert torch.equal(x, torch.cat([y, z], dim=1))
    for e in (whole, lo, hi):
        e.close()


def test_dict_api_drop_in_matches_oracle(sc, ph):
    """num_envs == 1: the reference's reset()/step() dict contract, compared with the
    object-level oracle stepping the same env definition."""
    seed = 31
    env = sc.SupplyChainEnv(seed=seed)
    st = wl.order_stream(seed, 0)
    ref = wl.build(po, st)
    clock = harness.EpisodeClock([st])
    assert env.agent_ids == ref.agent_ids and env.strategic_agent_ids == ["SHOP"]
    assert env.non_strategic_agent_ids == ref.non_strategic_agent_ids
    for ep in range(2):
        clock.on_reset()
        o_ref, i_ref = ref.reset()
        o, i = env.reset()
        assert list(o) == ["SHOP"] and i == {} and np.array_equal(o["SHOP"], o_ref["SHOP"])
        assert o["SHOP"].dtype == np.float32
        r = np.random.RandomState(ep)
        for t in range(100):
            a = {"SHOP": r.uniform(0, 100, size=(1,)).astype(np.float32)} if t % 7 else {}
            clock.on_step(ref)
            s_ref = ref.step(a)
            s = env.step(a)
            assert isinstance(s, ph.PhantomEnv.Step)
            assert np.array_equal(s.observations["SHOP"], s_ref.observations["SHOP"])
            # float rewards: within 1e-5 relative of the reference's float64 (spec tolerance)
            assert s.rewards["SHOP"] == pytest.approx(s_ref.rewards["SHOP"], rel=1e-5, abs=1e-7)
            assert s.terminations == s_ref.terminations and s.truncations == s_ref.truncations
            assert s.infos == s_ref.infos
            assert env.current_step == ref.current_step
            shop = env.agents["SHOP"]
            assert (shop.stock, shop.sales, shop.missed_sales) == wl.shop_state(ref)[:3]
    env.close()


def test_dict_api_tracked_messages(sc, ph):
    env = sc.SupplyChainEnv(seed=3, enable_tracking=True)
    env.reset()
    env.step({"SHOP": np.array([40.0], np.float32)})
    msgs = env.network.resolver.tracked_messages
    assert len(msgs) == 12
    assert msgs[0] == ph.Message("SHOP", "WAREHOUSE", sc.StockRequest(40))
    assert [m.sender_id for m in msgs[1:6]] == [f"CUST{i}" for i in range(1, 6)]
    assert msgs[6] == ph.Message("WAREHOUSE", "SHOP", sc.StockResponse(40))
    assert [m.receiver_id for m in msgs[7:]] == [f"CUST{i}" for i in range(1, 6)]
    assert all(isinstance(m.payload, sc.OrderResponse) for m in msgs[7:])
    env.network.resolver.clear_tracked_messages()
    assert env.network.resolver.tracked_messages == []
    env.close()


def test_faults_raise_reference_exceptions(sc, ph):
    # missing SHOP<->WAREHOUSE edge: NetworkError on the first send (network.py:246-249)
    env = sc.SupplyChainEnv()
    del env.network._succ["SHOP"]["WAREHOUSE"], env.network._succ["WAREHOUSE"]["SHOP"]
    env.reset()
    with pytest.raises(ph.NetworkError):
        env.step({"SHOP": np.array([1.0], np.float32)})
    env.close()
    # ... unless the shop sends nothing and the customers' edges exist
    env = sc.SupplyChainEnv()
    del env.network._succ["SHOP"]["WAREHOUSE"], env.network._succ["WAREHOUSE"]["SHOP"]
    env.reset()
    env.step({})
    env.close()
    # round_limit=1: responses remain queued -> RuntimeError (resolvers.py:160-163)
    env = sc.SupplyChainEnv()
    env.network.resolver.round_limit = 1
    env.reset()
    with pytest.raises(RuntimeError):
        env.step({"SHOP": np.array([1.0], np.float32)})
    env.close()
    # round_limit=2 is enough
    env = sc.SupplyChainEnv()
    env.network.resolver.round_limit = 2
    env.reset()
    env.step({"SHOP": np.array([1.0], np.float32)})
    env.close()
    # non-finite action: int(round(nan)) raises ValueError in the reference
    env = sc.SupplyChainEnv()
    env.reset()
    with pytest.raises(ValueError):
        env.step({"SHOP": np.array([np.nan], np.float32)})
    env.close()


def test_payload_whitelist_violation(sc, ph):
    """A customer subclass with a different class name is not in OrderRequest's sender
    whitelist ('CustomerAgent', exact class-name match, network.py:315-331)."""

    class VipCustomer(sc.CustomerAgent):
        pass

    env = sc.SupplyChainEnv()
    env.network.agents["CUST3"].__class__ = VipCustomer
    env.reset()
    with pytest.raises(ph.NetworkError):
        env.step({"SHOP": np.array([1.0], np.float32)})
    env.close()
    env = sc.SupplyChainEnv()
    env.network.agents["CUST3"].__class__ = VipCustomer
    env.network.enforce_msg_payload_checks = False
    env.reset()
    env.step({"SHOP": np.array([1.0], np.float32)})
    env.close()


def test_ignore_connection_errors_drops_undeliverable_mail(sc):
    """ignore_connection_errors=True: a send over a missing edge is accepted but filtered at
    delivery (resolvers.py:146-148) -- the disconnected customer's order is never served."""
    seed = 77
    env = sc.SupplyChainEnv(seed=seed)
    env.network.ignore_connection_errors = True
    del env.network._succ["SHOP"]["CUST2"], env.network._succ["CUST2"]["SHOP"]
    st = wl.order_stream(seed, 0)
    ref = wl.build(po, st)
    ref.network.ignore_connection_errors = True
    del ref.network.graph._succ["SHOP"]["CUST2"], ref.network.graph._succ["CUST2"]["SHOP"]
    clock = harness.EpisodeClock([st])
    clock.on_reset(); ref.reset(); env.reset()
    for t in range(30):
        a = {"SHOP": np.array([7.0 * (t % 5)], np.float32)}
        clock.on_step(ref)
        s_ref, s = ref.step(a), env.step(a)
        assert np.array_equal(s.observations["SHOP"], s_ref.observations["SHOP"]), t
    env.close()


def test_auto_reset_same_step(sc):
    """PHX_FLAG_AUTO_RESET: the step that ends an episode also resets the env; its obs row
    then holds the reset observation (stock 0, sales/missed carried over)."""
    E, seed, T = 64, 4, 250
    A = np.random.RandomState(3).uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
    env = sc.SupplyChainEnv(num_envs=E, seed=seed, auto_reset=True)
    v = vectorised.SupplyChainVec(E, seed)
    env.reset_batch(); v.reset()
    ro = env.rollout_batch(A)
    obs = ro.observations.cpu().numpy()[:, :, 0]
    rew = ro.rewards.cpu().numpy()[:, :, 0]
    ad = ro.all_done.cpu().numpy()
    for t in range(T):
        ref = v.step(A[t, :, 0, 0])
        want_obs = ref["obs"]
        if ref["all_trunc"].all():
            want_obs = v.reset()
        assert np.array_equal(obs[t], want_obs), t
        assert np.array_equal(rew[t], ref["reward"].astype(np.float32)), t
        assert np.array_equal(ad[t, :, 1], ref["all_trunc"]), t
    assert ad[:, :, 1].sum() == 2 * E
    env.close()


def test_rollout_host_roundtrip(sc):
    E, T, seed = 4096, 10, 8
    A = np.random.RandomState(4).uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
    a = sc.SupplyChainEnv(num_envs=E, seed=seed)
    b = sc.SupplyChainEnv(num_envs=E, seed=seed)
    a.reset_batch(); b.reset_batch()
    dev = a.rollout_batch(A)
    host = b.rollout_host(A)
    assert np.array_equal(host["observations"], dev.observations.cpu().numpy())
    assert np.array_equal(host["rewards"], dev.rewards.cpu().numpy())
    assert np.array_equal(host["all_done"], dev.all_done.cpu().numpy())
    a.close(); b.close()


@pytest.mark.parametrize("exec_mode", ["fast", "queue"])
def test_nondefault_customer_count(sc, exec_mode):
    """Runtime-N path of the fast kernel (N != 5) and the wider queue tiles (G = 16, 32)
    against the vectorised oracle."""
    for n in (1, 3, 8, 13):
        E, T, seed = 257, 25, 12
        A = np.random.RandomState(n).uniform(0, 100, size=(T, E, 1, 1)).astype(np.float32)
        env = sc.SupplyChainEnv(n, num_envs=E, seed=seed, exec_mode=exec_mode)
        v = vectorised.SupplyChainVec(E, seed, n_customers=n)
        env.reset_batch(); v.reset()
        ro = env.rollout_batch(A)
        obs = ro.observations.cpu().numpy()[:, :, 0]
        for t in range(T):
            ref = v.step(A[t, :, 0, 0])
            assert np.array_equal(obs[t], ref["obs"]), (n, t)
        env.close()


def test_ratio_exhaustive():
    """Observation arithmetic: the kernel's float32(n / d) (Markstein FMA sequence with the
    correctly rounded reciprocal) equals numpy's float32(float64(n) / d) -- the reference's
    `np.array([n / d], dtype=np.float32)` -- for EVERY n the action contract allows
    (|n| <= 2^21) and the denominators of the shipped configs."""
    from phantom_b200 import _lib as L

    # (100, 25: supply chain; 10, 15, 30: Stackelberg game C4; 12, 33: FSM market C3)
    for den in (100, 25, 5, 40, 65, 10, 15, 30, 12, 33):
        lo, count = -(1 << 21), (1 << 22) + 1
        out = np.empty(count, np.float32)
        L.check(L.lib.phx_selftest_ratio(0, den, lo, count, out.ctypes.data))
        n = np.arange(lo, lo + count, dtype=np.float64)
        want = (n / den).astype(np.float32)
        assert np.array_equal(out, want), den


def test_queue_engine_equals_fast_kernel_full_episode(sc):
    """Invariance under the kernel variant: schedule-specialised kernel == dynamic queue."""
    E, T, seed = 4096, 100, 77
    A = np.random.RandomState(5).uniform(-10, 130, size=(T, E, 1, 1)).astype(np.float32)
    M = (np.random.RandomState(6).uniform(size=(T, E, 1)) > 0.1).astype(np.uint8)
    f = sc.SupplyChainEnv(num_envs=E, seed=seed, exec_mode="fast")
    q = sc.SupplyChainEnv(num_envs=E, seed=seed, exec_mode="queue")
    th = sc.SupplyChainEnv(num_envs=E, seed=seed, exec_mode="thread")
    assert f.exec_name.startswith("fast") and q.exec_name.startswith("queue(G=8")
    assert th.exec_name.startswith("thread-per-env")
    f.reset_batch(); q.reset_batch(); th.reset_batch()
    a, b, c = f.rollout_batch(A, M), q.rollout_batch(A, M), th.rollout_batch(A, M)
    for x, y, z in zip(a, b, c):
        assert torch.equal(x, y) and torch.equal(x, z)
    assert np.array_equal(shop_state(f), shop_state(q))
    assert np.array_equal(shop_state(f), shop_state(th))
    f.close(); q.close(); th.close()


def test_queue_engine_any_agent_order(sc, ph):
    """The queue engine is not tied to the example's agent order: customers first, shop
    last.  Routing (receiver first-arrival order!) changes; compare with the object-level
    oracle running the same permuted env, including the tracked message order."""
    seed = 19

    def build(api, mods, stream=None):
        # agents in the order: CUST1, CUST2, WAREHOUSE, CUST3, SHOP
        if api is ph:
            agents = [mods.CustomerAgent("CUST1", "SHOP"), mods.CustomerAgent("CUST2", "SHOP"),
                      mods.FactoryAgent("WAREHOUSE"), mods.CustomerAgent("CUST3", "SHOP"),
                      mods.ShopAgent("SHOP", "WAREHOUSE")]
            net = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=True))
            net.add_connection("SHOP", "WAREHOUSE")
            net.add_connections_between(["SHOP"], ["CUST1", "CUST2", "CUST3"])
            env = ph.PhantomEnv(num_steps=20, network=net, seed=seed)
            env.max_order, env.max_stock = 5, 100
            return env
        return None

    env = build(ph, sc)
    assert env.exec_name.startswith("thread-per-env")
    # oracle twin with the same order
    st = wl.order_stream(seed, 0, n_customers=3)
    ref_full = wl.build(po, st, n_customers=3, num_steps=20, enable_tracking=True)
    order = ["CUST1", "CUST2", "WAREHOUSE", "CUST3", "SHOP"]
    ref_full.network.agents = {k: ref_full.network.agents[k] for k in order}
    clock = harness.EpisodeClock([st])
    clock.on_reset(); ref_full.reset(); env.reset()
    slot = {aid: i for i, aid in enumerate(order)}
    r = np.random.RandomState(1)
    for t in range(20):
        a = {"SHOP": r.uniform(0, 60, size=(1,)).astype(np.float32)}
        clock.on_step(ref_full)
        ref_full.network.resolver.clear_tracked_messages()
        env.network.resolver.clear_tracked_messages()
        s_ref, s = ref_full.step(a), env.step(a)
        assert np.array_equal(s.observations["SHOP"], s_ref.observations["SHOP"]), t
        assert s.truncations == s_ref.truncations
        got = [(m.sender_id, m.receiver_id, type(m.payload).__name__, m.payload.size)
               for m in env.network.resolver.tracked_messages]
        want = [(m.sender_id, m.receiver_id, type(m.payload).__name__, m.payload.size)
                for m in ref_full.network.resolver.tracked_messages]
        assert got == want, t
    env.close()
