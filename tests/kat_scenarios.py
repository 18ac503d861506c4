"""The reference's own step-loop / routing known-answer tests, restated once and run twice:
on the CPU against the oracle (tests/test_oracle_kats.py) and on the B200 against the CUDA
engine through phantom_b200 (tests/test_gpu_kats.py).

Every scenario cites the reference test it restates; the expected literals are the
reference's.  `K` is a namespace with the API module (`ph`), the mock agent classes and
`finish_network` (see oracle/workloads/mock.py, phantom_b200/envs/mock.py).
"""
import numpy as np
import pytest


def approx_obs(got, want):
    """Observations: the reference returns float64 arrays, the device float32 (spec tolerance
    1e-5 relative)."""
    assert set(got) == set(want), (got, want)
    for k in want:
        np.testing.assert_allclose(np.asarray(got[k], np.float64), np.asarray(want[k], np.float64),
                                   rtol=1e-5, atol=1e-7)


def counts(agent):
    return (int(agent.compute_reward_count), int(agent.encode_obs_count),
            int(agent.decode_action_count))


def scenario_env_step_with_done_dropout(K):
    """/root/reference/tests/test_env.py:9-21,71-107"""
    ph = K.ph
    env = ph.PhantomEnv(
        num_steps=2,
        network=K.finish_network(ph.Network([
            K.MockStrategicAgent("A", num_steps=1), K.MockStrategicAgent("B"), K.MockAgent("C")])),
    )
    assert env.n_agents == 3
    assert env.agent_ids == ["A", "B", "C"]
    assert env.strategic_agent_ids == ["A", "B"]
    assert env.non_strategic_agent_ids == ["C"]
    assert env.strategic_agents == [env.agents["A"], env.agents["B"]]
    assert env.non_strategic_agents == [env.agents["C"]]
    assert env["A"].id == "A"

    obs, infos = env.reset()
    assert env.current_step == 0
    assert list(obs.keys()) == ["A", "B"]
    assert infos == {}

    step = env.step({"A": 0, "B": 0})
    assert env.current_step == 1
    assert list(step.observations.keys()) == ["A", "B"]
    assert list(step.rewards.keys()) == ["A", "B"]
    assert list(step.infos.keys()) == ["A", "B"]
    assert step.terminations == {"A": True, "B": False, "__all__": False}
    assert step.truncations == {"A": True, "B": False, "__all__": False}

    step = env.step({"A": 0, "B": 0})
    assert env.current_step == 2
    assert list(step.observations.keys()) == ["B"]
    assert list(step.rewards.keys()) == ["B"]
    assert list(step.infos.keys()) == ["B"]
    assert step.terminations == {"B": False, "__all__": False}
    assert step.truncations == {"B": False, "__all__": True}
    return env


def scenario_is_terminated_truncated(K):
    """/root/reference/tests/test_env.py:44-68, driven through real steps instead of poking
    the private sets: all strategic agents done => __all__ terminated and truncated."""
    ph = K.ph
    env = ph.PhantomEnv(
        num_steps=5,
        network=K.finish_network(ph.Network([
            K.MockStrategicAgent("A", num_steps=1), K.MockStrategicAgent("B", num_steps=2),
            K.MockAgent("C")])),
    )
    env.reset()
    s = env.step({"A": 0, "B": 0})
    assert s.terminations == {"A": True, "B": False, "__all__": False}
    s = env.step({"B": 0})
    assert s.terminations == {"B": True, "__all__": True}
    assert s.truncations == {"B": True, "__all__": True}
    return env


def _fsm_one_agent(K):
    ph = K.ph
    network = K.finish_network(ph.Network([K.MockStrategicAgent("agent")]))
    return ph.FiniteStateMachineEnv(
        num_steps=3, network=network, initial_stage="ODD",
        stages=[
            ph.FSMStage(stage_id="ODD", acting_agents=["agent"], next_stages=["EVEN"]),
            ph.FSMStage(stage_id="EVEN", acting_agents=["agent"], next_stages=["ODD"]),
        ])


def scenario_fsm_odd_even_one_agent(K):
    """/root/reference/tests/fsm/test_odd_even_one_agent.py:40-61"""
    env = _fsm_one_agent(K)
    obs, info = env.reset()
    approx_obs(obs, {"agent": np.array([0])})
    assert info == {}
    assert env.current_stage == "ODD"
    assert counts(env.agents["agent"]) == (0, 1, 0)

    step = env.step({"agent": np.array([0])})
    assert env.current_stage == "EVEN"
    approx_obs(step.observations, {"agent": np.array([1.0 / 3.0])})
    assert step.rewards == {"agent": 0.0}
    assert step.terminations == {"agent": False, "__all__": False}
    assert step.truncations == {"agent": False, "__all__": False}
    assert step.infos == {"agent": {}}
    assert counts(env.agents["agent"]) == (1, 2, 1)
    return env


def scenario_fsm_odd_even_two_agents(K):
    """/root/reference/tests/fsm/test_odd_even_two_agents.py:42-104 (incl. the `None` reward of
    an agent that observes before it was ever rewarded)"""
    ph = K.ph
    network = K.finish_network(ph.Network(
        [K.MockStrategicAgent("odd_agent"), K.MockStrategicAgent("even_agent")]))
    env = ph.FiniteStateMachineEnv(
        num_steps=3, network=network, initial_stage="ODD",
        stages=[
            ph.FSMStage(stage_id="ODD", next_stages=["EVEN"], acting_agents=["odd_agent"],
                        rewarded_agents=["odd_agent"]),
            ph.FSMStage(stage_id="EVEN", next_stages=["ODD"], acting_agents=["even_agent"],
                        rewarded_agents=["even_agent"]),
        ])
    obs, info = env.reset()
    approx_obs(obs, {"odd_agent": np.array([0])})
    assert env.current_stage == "ODD"
    assert counts(env.agents["odd_agent"]) == (0, 1, 0)
    assert counts(env.agents["even_agent"]) == (0, 0, 0)

    step = env.step({"odd_agent": np.array([1])})
    assert env.current_stage == "EVEN"
    approx_obs(step.observations, {"even_agent": np.array([1.0 / 3.0])})
    assert step.rewards == {"even_agent": None}
    assert step.terminations == {"even_agent": False, "odd_agent": False, "__all__": False}
    assert step.truncations == {"even_agent": False, "odd_agent": False, "__all__": False}
    assert step.infos == {"even_agent": {}}
    assert counts(env.agents["odd_agent"]) == (1, 1, 1)
    assert counts(env.agents["even_agent"]) == (0, 1, 0)

    step = env.step({"even_agent": np.array([0])})
    assert env.current_stage == "ODD"
    approx_obs(step.observations, {"odd_agent": np.array([2.0 / 3.0])})
    assert step.rewards == {"odd_agent": 0.0}
    assert step.terminations == {"even_agent": False, "odd_agent": False, "__all__": False}
    assert step.truncations == {"even_agent": False, "odd_agent": False, "__all__": False}
    assert step.infos == {"odd_agent": {}}
    assert counts(env.agents["odd_agent"]) == (1, 2, 1)
    assert counts(env.agents["even_agent"]) == (1, 1, 1)

    # not in the reference test: the terminal step flushes the caches (fsm.py:360-375)
    step = env.step({"odd_agent": np.array([1])})
    assert step.truncations["__all__"] is True
    approx_obs(step.observations, {"odd_agent": np.array([2.0 / 3.0]), "even_agent": np.array([1.0])})
    assert step.rewards == {"odd_agent": 0.0, "even_agent": 0.0}
    return env


def scenario_fsm_one_state(K):
    """/root/reference/tests/fsm/test_one_state.py:88-147 (handler-less variant, self-loop)"""
    ph = K.ph
    network = K.finish_network(ph.Network([K.MockStrategicAgent("agent")]))
    network.add_connection("agent", "agent")
    env = ph.FiniteStateMachineEnv(
        num_steps=2, network=network, initial_stage="UNIT",
        stages=[ph.FSMStage(stage_id="UNIT", acting_agents=["agent"], next_stages=["UNIT"],
                            handler=None)])
    obs, info = env.reset()
    approx_obs(obs, {"agent": np.array([0.0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (0, 1, 0)

    step = env.step({"agent": np.array([0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (1, 2, 1)
    approx_obs(step.observations, {"agent": np.array([0.5])})
    assert step.rewards == {"agent": 0}
    assert step.terminations == {"agent": False, "__all__": False}
    assert step.truncations == {"agent": False, "__all__": False}
    assert step.infos == {"agent": {}}

    step = env.step({"agent": np.array([0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (2, 3, 2)
    approx_obs(step.observations, {"agent": np.array([1.0])})
    assert step.rewards == {"agent": 0}
    assert step.terminations == {"agent": False, "__all__": False}
    assert step.truncations == {"agent": False, "__all__": True}
    assert step.infos == {"agent": {}}
    return env


def scenario_fsm_validation(K):
    """/root/reference/tests/fsm/test_fsm_validation.py:7-174 and
    tests/fsm/test_is_fsm_deterministic.py:4-47 (the list-registration cases)"""
    ph = K.ph
    net = lambda: K.finish_network(ph.Network([]))
    with pytest.raises(ph.fsm.FSMValidationError):  # no stages
        ph.FiniteStateMachineEnv(num_steps=1, network=net(), initial_stage="A")
    with pytest.raises(ph.fsm.FSMValidationError):  # bad initial stage
        ph.FiniteStateMachineEnv(
            num_steps=1, network=net(), initial_stage="X",
            stages=[ph.FSMStage(stage_id="A", acting_agents=[], next_stages=["A"])])
    with pytest.raises(ph.fsm.FSMValidationError):  # bad next stage
        ph.FiniteStateMachineEnv(
            num_steps=1, network=net(), initial_stage="A",
            stages=[ph.FSMStage(stage_id="A", acting_agents=[], next_stages=["B"])])
    with pytest.raises(ph.fsm.FSMValidationError):  # handler-less stage needs exactly one next
        ph.FiniteStateMachineEnv(
            num_steps=1, network=net(), initial_stage="A",
            stages=[ph.FSMStage(stage_id="A", acting_agents=[], next_stages=[])])
    env = ph.FiniteStateMachineEnv(
        num_steps=1, network=net(), initial_stage="A",
        stages=[ph.FSMStage(stage_id="A", acting_agents=[], next_stages=["B"]),
                ph.FSMStage(stage_id="B", acting_agents=[], next_stages=["C"]),
                ph.FSMStage(stage_id="C", acting_agents=[], next_stages=["A"])])
    assert env.is_fsm_deterministic()


def scenario_fsm_one_state_with_handler(K):
    """/root/reference/tests/fsm/test_one_state.py:15-69: one stage registered through the
    FSMStage DECORATOR whose env handler returns the stage itself (and does not resolve the
    network).  `K.stage_handler` is a Python function on the oracle / reference and a
    `StageRule` on the device."""
    ph = K.ph

    class OneStateFSMEnvWithHandler(ph.FiniteStateMachineEnv):
        def __init__(self):
            network = K.finish_network(ph.Network([K.MockStrategicAgent("agent")]))
            network.add_connection("agent", "agent")
            super().__init__(num_steps=2, network=network, initial_stage="UNIT")

        handle = ph.FSMStage(stage_id="UNIT", acting_agents=["agent"], next_stages=["UNIT"])(
            K.stage_handler("UNIT", resolve_network=False))

    env = OneStateFSMEnvWithHandler()
    obs, info = env.reset()
    approx_obs(obs, {"agent": np.array([0.0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (0, 1, 0)

    step = env.step({"agent": np.array([0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (1, 2, 1)
    approx_obs(step.observations, {"agent": np.array([0.5])})
    assert step.rewards == {"agent": 0}
    assert step.terminations == {"agent": False, "__all__": False}
    assert step.truncations == {"agent": False, "__all__": False}
    assert step.infos == {"agent": {}}

    step = env.step({"agent": np.array([0])})
    assert env.current_stage == "UNIT"
    assert counts(env.agents["agent"]) == (2, 3, 2)
    approx_obs(step.observations, {"agent": np.array([1.0])})
    assert step.rewards == {"agent": 0}
    assert step.terminations == {"agent": False, "__all__": False}
    assert step.truncations == {"agent": False, "__all__": True}
    assert step.infos == {"agent": {}}
    return env


def scenario_fsm_invalid_transition_runtime(K):
    """/root/reference/tests/fsm/test_fsm_validation.py:145-174: a handler returning a stage
    outside the stage's next_stages raises FSMRuntimeError from step() (fsm.py:304-307); and
    tests/fsm/test_is_fsm_deterministic.py:27-47: stages with two next stages and handlers
    make the FSM non-deterministic."""
    ph = K.ph
    network = K.finish_network(ph.Network([K.MockStrategicAgent("agent")]))
    env = ph.FiniteStateMachineEnv(
        num_steps=1, network=network, initial_stage="StageA",
        stages=[ph.FSMStage(stage_id="StageA", acting_agents=["agent"], next_stages=["StageA"],
                            handler=K.stage_handler("StageB"))])
    env.reset()
    with pytest.raises(ph.fsm.FSMRuntimeError):
        env.step({"agent": np.array([0])})

    env2 = ph.FiniteStateMachineEnv(
        num_steps=1, network=K.finish_network(ph.Network([])), initial_stage="A",
        stages=[ph.FSMStage(stage_id="A", acting_agents=[], next_stages=["A", "B"],
                            handler=K.stage_handler("B")),
                ph.FSMStage(stage_id="B", acting_agents=[], next_stages=["A", "B"],
                            handler=K.stage_handler("A"))])
    assert not env2.is_fsm_deterministic()


def _fsm_state_driven(K, **kw):
    ph = K.ph
    agents = [K.MockStrategicAgent("s"), K.MockStrategicAgent("t"),
              K.EchoAgent("a", seed_value=4), K.EchoAgent("b")]
    network = ph.Network(agents)
    network.add_connection("a", "b")
    return ph.FiniteStateMachineEnv(
        num_steps=6, network=K.finish_network(network), initial_stage="FILL",
        stages=[
            # the handler resolves the network, then looks at what the resolution did to `b`
            ph.FSMStage(stage_id="FILL", acting_agents=["s", "a"], rewarded_agents=["s"],
                        next_stages=["FILL", "DRAIN"],
                        handler=K.stage_handler("DRAIN", ("agent", "b", "handled_count"), ">=", 4,
                                                otherwise="FILL")),
            # no messages in this stage: the handler only looks at the clock
            ph.FSMStage(stage_id="DRAIN", acting_agents=["t"], rewarded_agents=["t"],
                        next_stages=["FILL", "DRAIN"],
                        handler=K.stage_handler("FILL", "step", ">=", 4, otherwise="DRAIN",
                                                resolve_network=False)),
        ], **kw)


# (stage after the step, observations, rewards, truncations["__all__"], a.handled_count,
#  b.handled_count, b.handled_total) -- produced by the UNMODIFIED reference running this
# scenario (oracle/workloads/mock.py classes on phantom imported through oracle/ref_shim.py) and
# identically by the oracle port; DESIGN.md 3 records the run.
FSM_STATE_DRIVEN_TRACE = [
    ("FILL", {"s": 1 / 6}, {"s": 0.0}, False, 1, 2, 5),
    ("DRAIN", {"t": 2 / 6}, {"t": None}, False, 2, 4, 10),
    ("DRAIN", {"t": 3 / 6}, {"t": 0.0}, False, 2, 4, 10),
    ("FILL", {"s": 4 / 6}, {"s": 0.0}, False, 2, 4, 10),
    ("DRAIN", {"t": 5 / 6}, {"t": 0.0}, False, 3, 6, 15),
    ("FILL", {"s": 1.0, "t": 5 / 6}, {"s": 0.0, "t": 0.0}, True, 3, 6, 15),
]


def scenario_fsm_handler_state_driven(K):
    """Handler-driven (non-deterministic) transitions, fsm.py:294-307: the next stage depends on
    agent state AFTER the handler's own resolve_network() (three resolver rounds of echo
    messages), and the observing set `acting_agents[next_stage]` (fsm.py:319-320) follows it."""
    env = _fsm_state_driven(K)
    obs, _ = env.reset()
    approx_obs(obs, {"s": np.array([0.0])})
    assert env.current_stage == "FILL"
    for want_stage, want_obs, want_rew, all_trunc, ac, bc, bt in FSM_STATE_DRIVEN_TRACE:
        step = env.step({"s": np.array([0]), "t": np.array([0])})
        assert env.current_stage == want_stage
        approx_obs(step.observations, {k: np.array([v]) for k, v in want_obs.items()})
        assert dict(step.rewards) == want_rew
        assert step.terminations == {"s": False, "t": False, "__all__": False}
        assert step.truncations == {"s": False, "t": False, "__all__": all_trunc}
        assert int(env.agents["a"].handled_count) == ac
        assert int(env.agents["b"].handled_count) == bc
        assert int(env.agents["b"].handled_total) == bt
    return env


def scenario_stackelberg(K):
    """/root/reference/tests/test_stackelberg.py:10-74"""
    ph = K.ph
    network = K.finish_network(ph.Network(
        [K.MockStrategicAgent("leader"), K.MockStrategicAgent("follower")]))
    env = ph.StackelbergEnv(3, network, ["leader"], ["follower"])
    obs, info = env.reset()
    approx_obs(obs, {"leader": np.array([0])})
    assert info == {}
    assert counts(env.agents["leader"]) == (0, 1, 0)
    assert counts(env.agents["follower"]) == (0, 0, 0)

    step = env.step({"leader": np.array([0])})
    approx_obs(step.observations, {"follower": np.array([1 / 3])})
    assert step.rewards == {}
    assert step.terminations == {"leader": False, "follower": False, "__all__": False}
    assert step.truncations == {"leader": False, "follower": False, "__all__": False}
    assert step.infos == {"follower": {}}
    assert counts(env.agents["leader"]) == (1, 1, 1)
    assert counts(env.agents["follower"]) == (0, 1, 0)

    step = env.step({"follower": np.array([0])})
    approx_obs(step.observations, {"leader": np.array([2 / 3])})
    assert step.rewards == {"leader": 0.0}
    assert step.terminations == {"leader": False, "follower": False, "__all__": False}
    assert step.truncations == {"leader": False, "follower": False, "__all__": False}
    assert step.infos == {"leader": {}}
    assert counts(env.agents["leader"]) == (1, 2, 1)
    assert counts(env.agents["follower"]) == (1, 1, 1)

    step = env.step({"leader": np.array([0])})
    approx_obs(step.observations, {"follower": np.array([1])})
    assert step.rewards == {"leader": 0.0, "follower": 0.0}
    assert step.terminations == {"leader": False, "follower": False, "__all__": False}
    assert step.truncations == {"leader": False, "follower": False, "__all__": True}
    assert step.infos == {"follower": {}}
    assert counts(env.agents["leader"]) == (2, 2, 2)
    assert counts(env.agents["follower"]) == (1, 2, 1)
    return env


def _msgs(env):
    out = []
    for m in env.network.resolver.tracked_messages:
        p = m.payload
        out.append((m.sender_id, m.receiver_id, type(p).__name__,
                    getattr(p, "value", getattr(p, "cash", None))))
    return out


def scenario_tracking_golden_vector(K):
    """/root/reference/tests/network/test_tracking.py:29-57 -- THE routing golden vector: exact
    global message order over three rounds.  The test's two hand-made sends become A's
    generate_messages()."""
    ph = K.ph
    resolver = ph.resolvers.BatchResolver(enable_tracking=True)
    n = ph.Network([K.EchoAgent("A", seed_value=4), K.EchoAgent("B"), K.EchoAgent("C")], resolver)
    n.add_connection("A", "B")
    n.add_connection("A", "C")
    env = ph.PhantomEnv(num_steps=3, network=K.finish_network(n))
    env.reset()
    env.step({})
    assert _msgs(env) == [
        ("A", "B", "TestMessage", 4),
        ("A", "C", "TestMessage", 4),
        ("B", "A", "TestMessage", 2),
        ("C", "A", "TestMessage", 2),
        ("A", "B", "TestMessage", 1),
        ("A", "C", "TestMessage", 1),
    ]
    env.network.resolver.clear_tracked_messages()
    assert env.network.resolver.tracked_messages == []
    return env


def scenario_resolver_round_ordering(K):
    """/root/reference/tests/network/test_resolver.py:48-70: requests A->B, A->C, B->C; round 0
    receivers are visited B then C (first arrival), C answering A before B; in round 1 the
    responses are delivered B->A, C->A, C->B -- receivers A then B."""
    ph = K.ph
    n = ph.Network(
        [K.EchoAgent("A", seed_value=100, request_response=True),
         K.EchoAgent("B", seed_value=100, request_response=True), K.EchoAgent("C")],
        ph.resolvers.BatchResolver(enable_tracking=True))
    n.add_connection("A", "B")
    n.add_connection("A", "C")
    n.add_connection("B", "C")
    env = ph.PhantomEnv(num_steps=3, network=K.finish_network(n))
    env.reset()
    env.step({})
    assert _msgs(env) == [
        ("A", "B", "Request", 100), ("A", "C", "Request", 100), ("B", "C", "Request", 100),
        ("B", "A", "Response", 50), ("C", "A", "Response", 50), ("C", "B", "Response", 50),
    ]
    assert int(env.agents["A"].handled_count) == 2 and int(env.agents["C"].handled_count) == 2
    assert int(env.agents["B"].handled_total) == 150
    return env


def scenario_round_limit(K):
    """/root/reference/tests/network/test_resolver.py:73-86: round_limit=0 with a queued
    message raises (RuntimeError, resolvers.py:160-163)."""
    ph = K.ph
    n = ph.Network([K.EchoAgent("A", seed_value=1), K.EchoAgent("B")],
                   ph.resolvers.BatchResolver(round_limit=0))
    n.add_connection("A", "B")
    env = ph.PhantomEnv(num_steps=3, network=K.finish_network(n))
    env.reset()
    with pytest.raises(RuntimeError):
        env.step({})


def scenario_unknown_message_type(K):
    """/root/reference/tests/test_agent.py:55-68: a message for an agent without a handler for
    its payload type raises ValueError (agents.py:140-143)."""
    ph = K.ph
    n = ph.Network([K.EchoAgent("A", seed_value=3), K.MockAgent("B")])
    n.add_connection("A", "B")
    env = ph.PhantomEnv(num_steps=3, network=K.finish_network(n))
    env.reset()
    with pytest.raises(ValueError):
        env.step({})


def scenario_done_agent_mail_is_dropped(K):
    """resolvers.py:143-144 (SURVEY.md A.1 rule 5): mail for an agent without a context (done)
    is dropped silently -- no handler call, no error.  S has no handler for TestMessage, so a
    DELIVERED message would raise ValueError; S finishes in step 1, A only acts from step 2."""
    ph = K.ph
    n = ph.Network([K.EchoAgent("A", seed_value=6), K.MockStrategicAgent("S", num_steps=1)],
                   ph.resolvers.BatchResolver(enable_tracking=True))
    n.add_connection("A", "S")
    env = ph.FiniteStateMachineEnv(
        num_steps=4, network=K.finish_network(n), initial_stage="ONE",
        stages=[ph.FSMStage(stage_id="ONE", acting_agents=["S"], next_stages=["TWO"]),
                ph.FSMStage(stage_id="TWO", acting_agents=["A"], next_stages=["TWO"])])
    env.reset()
    s = env.step({"S": 0})
    assert s.terminations["S"] is True and _msgs(env) == []
    s = env.step({})
    assert _msgs(env) == [("A", "S", "TestMessage", 6)]
    assert "S" not in s.terminations
    return env


def scenario_codec_composition(K):
    """Encoder / decoder / reward-function composition on the step path
    (encoders.py:64-131, decoders.py:54-124, reward_functions.py:26-38; composition semantics
    pinned by /root/reference/tests/encoders/test_chained.py:23-43, test_dict.py:23-50,
    tests/decoders/test_chained.py:25-51, test_dict.py:27-54): sub-encoders are evaluated in
    list / dict order, Empty decoders produce no messages, Constant rewards."""
    ph = K.ph
    enc, dec, rf = ph.encoders, ph.decoders, ph.reward_functions
    a = K.CodecAgent(
        "A", enc.ChainedEncoder([enc.Constant((2,), 0.5), K.ElapsedTime()]).chain([enc.EmptyEncoder()]),
        dec.EmptyDecoder(), rf.Constant(1.5))
    b = K.CodecAgent(
        "B", enc.DictEncoder({"t": K.CurrentStep(), "z": enc.EmptyEncoder(), "c": enc.Constant((1,), -3.0)}),
        dec.ChainedDecoder([dec.EmptyDecoder(), dec.EmptyDecoder()]), rf.Constant(-2.0))
    assert len(a.observation_space.spaces) == 3 and set(b.observation_space.spaces) == {"t", "z", "c"}
    env = ph.PhantomEnv(num_steps=4, network=K.finish_network(ph.Network([a, b])))
    obs, _ = env.reset()

    def check(obs, step):
        assert isinstance(obs["A"], tuple) and len(obs["A"]) == 3
        np.testing.assert_allclose(obs["A"][0], [0.5, 0.5])
        np.testing.assert_allclose(obs["A"][1], [step / 4], rtol=1e-6)
        np.testing.assert_allclose(obs["A"][2], [0.0])
        assert list(obs["B"]) == ["t", "z", "c"]
        np.testing.assert_allclose(obs["B"]["t"], [float(step)])
        np.testing.assert_allclose(obs["B"]["z"], [0.0])
        np.testing.assert_allclose(obs["B"]["c"], [-3.0])

    check(obs, 0)
    for t in (1, 2, 3, 4):
        s = env.step({"A": np.array([0.0]), "B": (np.array([0.0]), np.array([0.0]))})
        check(s.observations, t)
        assert s.rewards == {"A": 1.5, "B": -2.0}
        assert s.truncations["__all__"] is (t == 4)
    # composition bookkeeping of the reference: chain() flattens, reset() propagates
    ce = enc.ChainedEncoder([enc.EmptyEncoder(), [enc.EmptyEncoder(), enc.EmptyEncoder()]])
    assert len(ce.encoders) == 3
    ce.reset()
    return env


def scenario_stackelberg_acting_order(K):
    """StackelbergEnv hands `leader_agents` to _handle_acting_agents as the user wrote it
    (/root/reference/phantom/stackelberg.py:133-140, env.py:320-336): the leaders act in LIST
    order, not in network order, and that is the push order of their messages.  e1 (listed
    first) and e0 both message e2; e2's float32 `level` (level * 0.5 + value per handled message)
    tells the two orders apart: [9, 3] then the reply 2 -> 5.75; network order would give 7.25."""
    ph = K.ph
    agents = [K.EchoAgent("e0", seed_value=3), K.EchoAgent("e1", seed_value=9), K.EchoAgent("e2"),
              K.MockStrategicAgent("lead"), K.MockStrategicAgent("follow")]
    network = ph.Network(agents)
    network.add_connection("e0", "e2")
    network.add_connection("e1", "e2")
    env = ph.StackelbergEnv(4, K.finish_network(network), ["e1", "lead", "e0"], ["follow"])
    env.reset()
    env.step({"lead": np.array([0])})
    levels = [float(np.asarray(env.agents[e].level).reshape(-1)[0]) for e in ("e0", "e1", "e2")]
    assert levels == [1.0, 3.0, 5.75], levels
    assert [int(np.asarray(env.agents[e].handled_count).reshape(-1)[0]) for e in ("e0", "e1", "e2")] == [1, 2, 3]
    return env


ALL = [
    scenario_env_step_with_done_dropout,
    scenario_is_terminated_truncated,
    scenario_fsm_odd_even_one_agent,
    scenario_fsm_odd_even_two_agents,
    scenario_fsm_one_state,
    scenario_fsm_validation,
    scenario_fsm_one_state_with_handler,
    scenario_fsm_invalid_transition_runtime,
    scenario_fsm_handler_state_driven,
    scenario_stackelberg,
    scenario_stackelberg_acting_order,
    scenario_tracking_golden_vector,
    scenario_resolver_round_ordering,
    scenario_round_limit,
    scenario_unknown_message_type,
    scenario_done_agent_mail_is_dropped,
    scenario_codec_composition,
]


# ---------------------------------------------------------------- random handler-driven FSMs
def random_handler_fsm(K, case_seed, compound=False, wide=False, floats=False, waiting=False, **kw):
    """A random FiniteStateMachineEnv over the mock agents: 1-4 stages with random acting /
    rewarded sets, handler-less stages and stages with env handlers (`K.stage_handler`: always /
    clock / echo-agent counters after the handler's own resolve_network()), next_stages that
    sometimes do NOT contain what the handler returns (FSMRuntimeError at run time), strategic
    agents that terminate mid-episode, echo agents exchanging halving messages.  Deterministic in
    `case_seed`; used as a differential fuzz oracle-vs-reference (CPU) and device-vs-oracle (GPU).
    compound=True: most handlers are if / elif / else chains of up to four branches, each one or
    two comparisons between the clock, echo-agent counters and constants (a different draw
    sequence, its own golden).
    wide=True: 33..120 agents (12-49 strategic, 21-70 echo agents on a sparse random graph) -- env
    classes wider than a warp, which run on the 128-lane block engine; compound handlers.
    floats=True: compound handlers whose comparisons may also be FLOAT32 ones -- an echo agent's
    `level` (a float32 recurrence over the handled values) against a float constant or another
    agent's level.
    waiting=True: float32 cases whose handlers may NOT call resolve_network() although the acting
    agents sent mail -- the mail waits in the resolver for a later step's resolve
    (fsm.py:280-283), together with whatever is pushed in between."""
    r = np.random.RandomState(case_seed + (4000 if waiting else 3000 if floats else 2000 if wide
                                           else 1000 if compound else 0))
    floats = floats or waiting
    ph = K.ph
    strat = [f"s{i}" for i in range(int(r.randint(12, 50) if wide else r.randint(1, 4)))]
    echo = [f"e{i}" for i in range(int(r.randint(21, 71) if wide else r.randint(0, 4)))]
    compound = compound or wide or floats
    if floats and not echo:
        echo = ["e0"]
    seeds = {e: int(r.choice([0, 0, 3, 4, 9])) for e in echo}
    agents = [K.MockStrategicAgent(a, num_steps=(int(r.randint(1, 7)) if r.uniform() < 0.3 else None))
              for a in strat]
    agents += [K.EchoAgent(e, seed_value=seeds[e]) for e in echo]
    agents = [agents[i] for i in r.permutation(len(agents))]
    network = ph.Network(agents)
    for i in range(len(echo)):
        for j in range(i + 1, len(echo)):
            if r.uniform() < (0.05 if wide else 0.6):
                network.add_connection(echo[i], echo[j])
    sids = [f"S{k}" for k in range(int(r.randint(1, 5)))]
    stages = []
    for sid in sids:
        acting = [a for a in strat + echo if r.uniform() < 0.6]
        rewarded = None if r.uniform() < 0.3 else [a for a in strat if r.uniform() < 0.6]
        kinds = ["none", "always", "step"] + (["agent"] if echo else [])
        kind = kinds[int(r.randint(len(kinds)))]
        then, otherwise = (sids[int(r.randint(len(sids)))] for _ in range(2))
        if kind == "none":
            stages.append(ph.FSMStage(stage_id=sid, acting_agents=acting, rewarded_agents=rewarded,
                                      next_stages=[then]))
            continue
        sends = any(seeds.get(a, 0) > 0 for a in acting)
        resolve = True if (sends and not waiting) else bool(r.uniform() < 0.5)
        cmp = ["<", "<=", "==", "!=", ">=", ">"][int(r.randint(6))]
        if compound and kind != "always" and r.uniform() < 0.75:
            def operand(rhs):
                kinds_ = ["step"] + (["agent"] if echo else []) + (["const", "const"] if rhs else [])
                k = kinds_[int(r.randint(len(kinds_)))]
                if k == "step":
                    return "step"
                if k == "const":
                    return int(r.randint(0, 12))
                return ("agent", echo[int(r.randint(len(echo)))],
                        ["handled_count", "handled_total"][int(r.randint(2))])

            def term():
                if floats and echo and r.uniform() < 0.5:  # a float32 comparison
                    lvl = lambda: ("agent", echo[int(r.randint(len(echo)))], "level")
                    consts = [0.5, 1.0, 1.5, 2.25, 3.0, 4.5, 6.75, 9.0]
                    rhs = lvl() if r.uniform() < 0.3 else consts[int(r.randint(len(consts)))]
                    return (lvl(), ["<", "<=", "==", "!=", ">=", ">"][int(r.randint(6))], rhs)
                return (operand(False), ["<", "<=", "==", "!=", ">=", ">"][int(r.randint(6))],
                        operand(True))

            first = term()
            also = term() if r.uniform() < 0.5 else None
            elifs = []
            for _ in range(int(r.randint(0, 4))):
                terms = [term(), term()] if r.uniform() < 0.4 else term()
                elifs.append((terms, sids[int(r.randint(len(sids)))]))
            handler = K.stage_handler(then, first[0], first[1], first[2], otherwise=otherwise,
                                      resolve_network=resolve, also=also, elifs=elifs)
            returned = [then] + [st for _, st in elifs] + [otherwise]
        elif kind == "always":
            handler = K.stage_handler(then, resolve_network=resolve)
            returned = [then]
        elif kind == "step":
            handler = K.stage_handler(then, "step", cmp, int(r.randint(1, 7)), otherwise=otherwise,
                                      resolve_network=resolve)
            returned = [then, otherwise]
        else:
            column = ["handled_count", "handled_total"][int(r.randint(2))]
            handler = K.stage_handler(then, ("agent", echo[int(r.randint(len(echo)))], column), cmp,
                                      int(r.randint(0, 30)), otherwise=otherwise,
                                      resolve_network=resolve)
            returned = [then, otherwise]
        allowed = [s for s in sids if s in returned or r.uniform() < 0.3]
        if r.uniform() < 0.1:  # a handler that may return a stage outside next_stages
            allowed = [s for s in allowed if s != returned[-1]] or [sids[0]]
        stages.append(ph.FSMStage(stage_id=sid, acting_agents=acting, rewarded_agents=rewarded,
                                  next_stages=allowed, handler=handler))
    env = ph.FiniteStateMachineEnv(num_steps=8, network=K.finish_network(network),
                                   initial_stage=sids[int(r.randint(len(sids)))], stages=stages, **kw)
    return env, strat, echo


def run_random_handler_fsm(K, case_seed, compound=False, prepare=None, wide=False, floats=False,
                           waiting=False):
    """Steps the random FSM to the end of its episode (or its first exception) and returns a
    plain-Python trace that is comparable across implementations."""
    env, strat, echo = random_handler_fsm(K, case_seed, compound=compound, wide=wide, floats=floats,
                                          waiting=waiting)
    floats = floats or waiting
    if prepare is not None:
        prepare(env)

    def plain(d):
        return {k: (None if v is None else
                    [round(float(x), 6) for x in np.asarray(v, np.float64).reshape(-1)])
                for k, v in d.items()}

    obs, _ = env.reset()
    trace = [("reset", str(env.current_stage), plain(obs))]
    for t in range(8):
        try:
            step = env.step({a: np.array([0]) for a in strat})
        except Exception as exc:  # the implementations must agree on the exception type too
            trace.append(("raise", type(exc).__name__))
            break
        trace.append((
            "step", str(env.current_stage), plain(step.observations), plain(step.rewards),
            {k: bool(v) for k, v in step.terminations.items()},
            {k: bool(v) for k, v in step.truncations.items()},
            [list(map(int, counts(env.agents[a]))) for a in strat],
            [[int(env.agents[e].handled_count), int(env.agents[e].handled_total)] +
             ([float(env.agents[e].level)] if floats else []) for e in echo]))
        if step.terminations["__all__"] or step.truncations["__all__"]:
            break
    if hasattr(env, "close"):
        env.close()
    return trace


# ---- random PhantomEnv / StackelbergEnv env classes over the mock agents (differential fuzz of the
# base and Stackelberg step loops incl. message tracking, round limits, bad edges and shuffled
# batches; tools/fuzz_campaign3.py ran thousands of them, a fixture of the unmodified reference
# pins 60: tests/golden/mock_env_fuzz_reference.json)
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
    """Steps one random env class to the end of its episode (or its first exception) and returns a
    plain-Python trace: observations, rewards, done flags, call counters, float32 levels and the
    tracked message list of every step.  The oracle / reference side shuffles batches with the
    contract's Fisher-Yates (oracle/harness.py patched_np_shuffle)."""
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


