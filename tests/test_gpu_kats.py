"""The reference's KATs (restated in tests/kat_scenarios.py) replayed through the CUDA engine:
phantom_b200's plugin API -> lowering -> C ABI -> generic queue engine, one env per handle."""
import pytest

pytestmark = pytest.mark.gpu

from . import kat_scenarios as kats  # noqa: E402


@pytest.fixture(scope="module")
def K():
    import phantom_b200 as ph
    from phantom_b200.envs import mock

    class NS:
        pass

    ns = NS()
    ns.ph = ph
    ns.MockAgent, ns.MockStrategicAgent, ns.EchoAgent = mock.MockAgent, mock.MockStrategicAgent, mock.EchoAgent
    ns.finish_network = lambda network: network
    ns.CodecAgent = mock.CodecAgent
    ns.ElapsedTime, ns.CurrentStep = ph.encoders.ElapsedTime, ph.encoders.CurrentStep
    ns.stage_handler = ph.StageRule  # the device form of an env stage handler
    return ns


@pytest.mark.parametrize("exec_mode", ["thread", "queue", "wide"])
@pytest.mark.parametrize("scenario", kats.ALL, ids=lambda f: f.__name__)
def test_kat_on_device(K, scenario, exec_mode, monkeypatch):
    """Every KAT on all three variants of the generic engine: one thread per env, a tile of
    lanes per env, and a 128-lane block per env (csrc/phx_engine_wide.cuh, forced here: it is
    what env classes of 33..128 agents run on)."""
    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", exec_mode)
    env = scenario(K)
    if env is not None:
        assert env.exec_name.startswith({"thread": "thread-per-env", "queue": "queue(G=",
                                         "wide": "wide(G=128)"}[exec_mode])
        env.close()


def test_batched_kat_every_env_identical(K):
    """The same FSM KAT for 3 000 envs at once: every env must reproduce the single-env trace
    (mock agents draw no random numbers)."""
    import numpy as np

    ph = K.ph
    network = ph.Network([K.MockStrategicAgent("odd_agent"), K.MockStrategicAgent("even_agent")])
    env = ph.FiniteStateMachineEnv(
        num_steps=3, network=network, initial_stage="ODD", num_envs=3000,
        stages=[
            ph.FSMStage(stage_id="ODD", next_stages=["EVEN"], acting_agents=["odd_agent"],
                        rewarded_agents=["odd_agent"]),
            ph.FSMStage(stage_id="EVEN", next_stages=["ODD"], acting_agents=["even_agent"],
                        rewarded_agents=["even_agent"]),
        ])
    obs, mask = env.reset_batch()
    assert mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    acts = np.zeros((3000, 2, 1), np.float32)
    out = env.step_batch(acts, np.array([[1, 0]] * 3000, np.uint8))
    assert out.obs_mask.cpu().numpy().tolist() == [[0, 1]] * 3000
    assert out.reward_mask.cpu().numpy().tolist() == [[0, 2]] * 3000  # even_agent: None
    np.testing.assert_allclose(out.observations.cpu().numpy()[:, 1, 0], 1 / 3, rtol=1e-6)
    out = env.step_batch(acts, np.array([[0, 1]] * 3000, np.uint8))
    assert out.obs_mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    assert out.reward_mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    out = env.step_batch(acts, np.array([[1, 0]] * 3000, np.uint8))
    assert out.all_done.cpu().numpy().tolist() == [[0, 1]] * 3000
    assert out.obs_mask.cpu().numpy().tolist() == [[1, 1]] * 3000  # terminal flush of the caches
    assert out.reward_mask.cpu().numpy().tolist() == [[1, 1]] * 3000
    env.close()


@pytest.mark.parametrize("exec_mode", ["thread", "queue"])
def test_handler_driven_fsm_envs_diverge(K, exec_mode, monkeypatch):
    """Handler-driven transitions (fsm.py:294-307) across a batch whose envs take DIFFERENT
    paths through the FSM: env e starts with b.handled_count = e % 7, so the FILL stage's rule
    (`b.handled_count >= 4` after the handler's resolve_network()) fires at a different step per
    env.  Stage column, observation / reward masks and the echo counters are compared per env
    with the oracle stepping one env object per start value."""
    import numpy as np

    import oracle.phantom_oracle as po
    from oracle.workloads import mock as omock

    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", exec_mode)
    E, T = 70, 6
    env = kats._fsm_state_driven(K, num_envs=E)
    env.reset_batch()
    start = np.arange(E, dtype=np.int32) % 7
    env.agents["b"].handled_count = start
    stage_ids = ["FILL", "DRAIN"]
    got_stage, got_om, got_rm, got_bc = [], [], [], []
    acts = np.zeros((E, 2, 1), np.float32)
    for _ in range(T):
        out = env.step_batch(acts)
        got_stage.append(np.asarray(env.current_stage).copy())
        got_om.append(out.obs_mask.cpu().numpy().copy())
        got_rm.append(out.reward_mask.cpu().numpy().copy())
        got_bc.append(np.asarray(env.agents["b"].handled_count).copy())
    env.close()

    KO = omock.build_classes(po)
    for k in range(7):
        ref = kats._fsm_state_driven(KO)
        ref.reset()
        ref.agents["b"].handled_count = int(k)
        for t in range(T):
            step = ref.step({"s": np.array([0]), "t": np.array([0])})
            rows = np.nonzero(start == k)[0]
            assert (got_stage[t][rows] == stage_ids.index(ref.current_stage)).all(), (k, t)
            want_om = [int(a in step.observations) for a in ("s", "t")]
            assert (got_om[t][rows] == want_om).all(), (k, t)
            # reward mask: 1 = a float reward, 2 = the key is present with value None
            want_rm = [0 if a not in step.rewards else (2 if step.rewards[a] is None else 1)
                       for a in ("s", "t")]
            assert (got_rm[t][rows] == want_rm).all(), (k, t)
            assert (got_bc[t][rows] == ref.agents["b"].handled_count).all(), (k, t)


def test_handler_without_resolve_leaves_mail(K):
    """A stage handler that never calls resolve_network() while agents sent messages: the
    reference keeps that mail queued for some later resolve.  The thread-per-env engine (the
    default for env classes of <= 8 agents) does too -- nothing is delivered, nothing is lost,
    the queue grows until it overflows loudly; the tile engine refuses at the first such step
    (sticky fault, RuntimeError) instead of silently dropping the mail."""
    import numpy as np

    ph = K.ph

    def build(**kw):
        network = ph.Network([K.MockStrategicAgent("s"), K.EchoAgent("a", seed_value=4), K.EchoAgent("b")])
        network.add_connection("a", "b")
        return ph.FiniteStateMachineEnv(
            num_steps=3, network=network, initial_stage="X",
            stages=[ph.FSMStage(stage_id="X", acting_agents=["s", "a"], next_stages=["X"],
                                handler=ph.StageRule("X", resolve_network=False))], **kw)

    env = build()
    assert env.exec_name.startswith("thread-per-env")
    env.reset()
    for _ in range(3):
        env.step({"s": np.array([0])})
    assert int(env.agents["b"].handled_count) == 0  # three messages wait, none was delivered
    env.close()
    env = build(exec_mode="queue")
    env.reset()
    with pytest.raises(RuntimeError, match="unresolved"):
        env.step({"s": np.array([0])})
    env.close()


def test_python_stage_handler_is_not_lowerable(K):
    ph = K.ph
    network = ph.Network([K.MockStrategicAgent("s")])
    env = ph.FiniteStateMachineEnv(
        num_steps=3, network=network, initial_stage="X",
        stages=[ph.FSMStage(stage_id="X", acting_agents=["s"], next_stages=["X"],
                            handler=lambda e: "X")])
    with pytest.raises(ph.NotLowerableError):
        env.reset()


@pytest.mark.parametrize("exec_mode", ["thread", "queue", "wide"])
def test_random_handler_fsms_match_the_reference(K, exec_mode, monkeypatch):
    """Differential fuzz of handler-driven FSM transitions on the device: 40 random
    FiniteStateMachineEnvs (tests/kat_scenarios.py:random_handler_fsm -- random stage tables,
    StageRules on the clock / echo-agent counters, invalid transitions, agents terminating
    mid-episode) == the traces of the UNMODIFIED reference running the same case seeds with
    Python handlers (tests/golden/fsm_handler_fuzz_reference.json)."""
    import json
    import os

    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", exec_mode)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "fsm_handler_fuzz_reference.json")
    want = json.load(open(path))
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s)))
        assert got == want[str(s)], f"case seed {s}"
    # compound StageRules: if / elif / else chains of one or two comparisons per branch, operands
    # on both sides (clock, echo-agent counters, constants)
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_compound_fuzz")))
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, compound=True)))
        assert got == want[str(s)], f"compound case seed {s}"
    # ... with FLOAT32 comparisons in the chains: an echo agent's float32 `level` (a recurrence
    # over the handled values, one rounding per operation) against float constants / other levels
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_float_fuzz")))
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, floats=True)))
        assert got == want[str(s)], f"float case seed {s}"
    # ... and the cases of a 2 500-seed campaign whose outcome depends on the ORDER of
    # FSMStage.acting_agents: the agents act in the user's list order (fsm.py:276-277), which
    # decides the order of a receiver's batch and with it an order-sensitive float32 recurrence
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_order_fuzz")))
    for s in want:
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, int(s), floats=True)))
        assert got == want[s], f"order case seed {s}"


def test_compound_stage_rules_specialised(K, monkeypatch):
    """The same compound-handler cases through run-time specialised units of the thread-per-env
    engine (the rule chain is a compile-time constant there and folds)."""
    import json
    import os

    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", "thread")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "fsm_compound_fuzz_reference.json")
    want = json.load(open(path))

    def prepare(env):
        env.specialise()
        assert "specialised" in env.exec_name

    for s in (0, 3, 4, 7, 11, 19, 34):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, compound=True, prepare=prepare)))
        assert got == want[str(s)], f"compound case seed {s}"
    want = json.load(open(path.replace("fsm_compound_fuzz", "fsm_float_fuzz")))
    for s in (1, 2, 5, 9, 17):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, floats=True, prepare=prepare)))
        assert got == want[str(s)], f"float case seed {s}"


def test_wide_random_fsms_match_the_reference(K):
    """16 random FSM env classes of 33..120 agents (12-49 strategic agents, 21-70 echo agents on a
    sparse graph, compound stage handlers, invalid transitions, agents terminating mid-episode)
    on the 128-lane block engine == the traces of the UNMODIFIED reference
    (tests/golden/fsm_wide_fuzz_reference.json)."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "fsm_wide_fuzz_reference.json")
    want = json.load(open(path))
    seen = []

    def prepare(env):
        env._ensure_handle()
        seen.append(env.exec_name)

    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, wide=True, prepare=prepare)))
        assert got == want[str(s)], f"wide case seed {s}"
    assert seen and all(n == "wide(G=128)" for n in seen)


def test_mail_that_waits_across_steps(K, monkeypatch):
    """A stage handler that does not call resolve_network() leaves the step's mail in the
    resolver; it is delivered -- behind it whatever was pushed in between -- by the next resolve
    (fsm.py:280-283).  The thread-per-env engine keeps that queue between steps and launches
    (EngineArgs.carry): 40 random FSMs == the traces of the UNMODIFIED reference
    (tests/golden/fsm_waiting_fuzz_reference.json), generic and specialised builds.  The tile and
    block engines still refuse loudly (PHX_FAULT_UNRESOLVED_MAIL)."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "fsm_waiting_fuzz_reference.json")
    want = json.load(open(path))
    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", "thread")
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, waiting=True)))
        assert got == want[str(s)], f"waiting case seed {s}"

    def prepare(env):
        env.specialise()

    for s in (0, 1, 2, 3, 5, 8, 13, 21):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, waiting=True, prepare=prepare)))
        assert got == want[str(s)], f"waiting case seed {s} (specialised)"
    # a case in which mail really waits, on the tile engine: refused, not mis-delivered
    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", "queue")
    differ = [s for s in range(len(want))
              if json.loads(json.dumps(kats.run_random_handler_fsm(K, s, waiting=True))) != want[str(s)]]
    assert differ, "no case exercised waiting mail"
    for s in differ:
        got = kats.run_random_handler_fsm(K, s, waiting=True)
        assert got[-1][0] == "raise", (s, got[-1])
    # inside ONE launch the waiting mail stays in shared memory between the steps, between
    # launches it goes through HBM: a T-step rollout == T single-step launches (auto-reset on)
    import numpy as np

    from .generic_parity import assert_batchsteps_equal

    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", "thread")
    checked = 0
    for s in differ[:6]:
        if want[str(s)][-1][0] == "raise":
            continue
        envs = [kats.random_handler_fsm(K, s, waiting=True, num_envs=64, seed=3, auto_reset=True)[0]
                for _ in range(2)]
        S = len(envs[0].strategic_agents)
        A = np.zeros((20, 64, S, 1), np.float32)
        for e in envs:
            e.reset_batch()
        ro = envs[0].rollout_batch(A)
        for t in range(20):
            out = envs[1].step_batch(A[t])
            assert_batchsteps_equal(type(out)(*[x[t] for x in ro]), out, f"case {s} step {t}")
        for e in envs:
            e.check_errors()
            e.close()
        checked += 1
    assert checked >= 2


@pytest.mark.parametrize("exec_mode", ["thread", "queue", "wide"])
def test_random_base_and_stackelberg_envs_match_the_reference(K, exec_mode, monkeypatch):
    """The 60 random PhantomEnv / StackelbergEnv env classes of
    tests/golden/mock_env_fuzz_reference.json (traces of the UNMODIFIED reference: observations,
    rewards, done flags, counters, float32 levels, exception types and the tracked message list of
    every step) on every engine tiling; the shuffled cases need the per-receiver batch lists of
    the tile / block engines."""
    import json
    import os

    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", exec_mode)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "mock_env_fuzz_reference.json")
    want = json.load(open(path))
    ran = 0
    for s in range(len(want)):
        if exec_mode == "thread" and want[str(s)][0][1]:
            continue
        got = json.loads(json.dumps(kats.run_mock_env(K, s)))
        assert got == want[str(s)], f"case seed {s}"
        ran += 1
    assert ran >= 30
