"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/phx.h declares, the ctypes mirror of phx_spec matches the compiled layout, lowering
produces the expected flat tables, and the product refuses to run without a GPU (no CPU
fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import phantom_b200 as ph
from phantom_b200 import _lib as L


def header_functions():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(here, "include", "phx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(phx_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    declared = header_functions()
    assert len(declared) >= 19
    lib = C.CDLL(L.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/phx.h but not exported"
    assert sorted(L.SYMBOLS) == declared, "ctypes binding and header disagree"


def test_spec_layout_matches_compiled_struct():
    assert L.lib.phx_sizeof_spec() == C.sizeof(L.PhxSpec)
    assert L.lib.phx_abi_version() == L.PHX_ABI_VERSION == 7


def test_no_cpu_fallback():
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    if L.lib.phx_device_count() > 0:
        pytest.skip("a GPU is present")
    env = SupplyChainEnv(num_envs=4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        env.reset_batch()
    # and the raw ABI refuses too
    h = C.c_void_p()
    spec = env.spec
    rc = L.lib.phx_create(C.byref(spec), 4, 0, 0, 0, C.byref(h))
    assert rc == L.PHX_ERR_NO_DEVICE and b"no CPU fallback" in L.lib.phx_last_error()


def test_bad_spec_is_rejected_before_touching_the_device():
    spec = L.PhxSpec()
    h = C.c_void_p()
    assert L.lib.phx_create(C.byref(spec), 4, 0, 0, 0, C.byref(h)) == L.PHX_ERR_INVALID
    assert b"struct_size" in L.lib.phx_last_error()
    assert L.lib.phx_create(None, 4, 0, 0, 0, C.byref(h)) == L.PHX_ERR_INVALID


def test_lowering_supply_chain_tables():
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    env = SupplyChainEnv()
    s = env.spec
    assert (s.family, s.env_kind, s.n_agents, s.n_strategic) == (L.FAMILY_SUPPLY_CHAIN, L.ENV_BASE, 7, 1)
    assert list(s.agent_kind[:7]) == [0, 1, 2, 2, 2, 2, 2]
    assert list(s.strategic_index[:7]) == [0, -1, -1, -1, -1, -1, -1]
    # star on SHOP, every connection in both directions (network.py:122-123)
    assert s.adjacency[0][0] == 0b1111110 and all(s.adjacency[i][0] == 1 for i in range(1, 7))
    # OrderRequest: CustomerAgent -> ShopAgent (exact class-name whitelists, network.py:315-331)
    assert s.type_sender_ok[0][0] == 0b1111100 and s.type_receiver_ok[0][0] == 0b1
    assert s.round_limit == -1 and s.num_steps == 100 and (s.obs_dim, s.act_dim) == (3, 1)
    assert (s.iparams[0], s.iparams[1]) == (5, 100)


def test_lowering_fsm_and_stackelberg_tables():
    from phantom_b200.envs.market import MarketEnv
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    s = MarketEnv().spec
    assert (s.env_kind, s.n_stages, s.initial_stage, s.n_agents, s.n_strategic) == (L.ENV_FSM, 3, 0, 32, 31)
    assert s.stages[0].acting[0] == 0x7F and s.stages[1].acting[0] == 0x7FFFFF80
    assert s.stages[2].acting[0] == 1 << 31 and s.stages[2].rewarded[0] == 0x7FFFFFFF
    assert [s.stages[k].next_stage for k in range(3)] == [1, 2, 0]
    assert [s.stages[k].rewarded_is_none for k in range(3)] == [0, 0, 0]
    s = StackelbergGameEnv().spec
    assert (s.env_kind, s.leaders[0], s.followers[0]) == (L.ENV_STACKELBERG, 0b0001, 0b1110)


def test_lowering_fsm_stage_handlers():
    """FSMStage.handler (fsm.py:294-307): a StageRule lowers to phx_stage.rule_*, a Python
    callable is refused, a rule returning an unknown stage gets an index outside next_allowed."""
    import phantom_b200 as ph
    from phantom_b200.envs import mock

    class NS:
        pass

    K = NS()
    K.ph, K.MockStrategicAgent, K.EchoAgent = ph, mock.MockStrategicAgent, mock.EchoAgent
    K.finish_network, K.stage_handler = (lambda n: n), ph.StageRule
    from . import kat_scenarios as kats

    s = kats._fsm_state_driven(K).spec
    fill, drain = s.stages[0], s.stages[1]
    assert (fill.handler, fill.rule_resolves, fill.next_allowed) == (2, 1, 0b11)
    assert (fill.rule_n_branches, fill.rule_else) == (1, 0)
    t = fill.rule_branch[0].term[0]
    assert (fill.rule_branch[0].n_terms, fill.rule_branch[0].then) == (1, 1)
    assert (t.lhs, t.slot, t.word) == (L.RULE_AGENT_WORD, 3, 3)  # b.handled_count
    assert (t.cmp, t.rhs_kind, t.rhs) == (L.CMP_GE, L.RULE_CONST, 4)
    t = drain.rule_branch[0].term[0]
    assert (drain.handler, drain.rule_resolves, t.lhs) == (2, 0, L.RULE_STEP)
    assert (t.cmp, t.rhs, drain.rule_branch[0].then, drain.rule_else) == (L.CMP_GE, 4, 0, 1)

    def one_stage(handler):
        return ph.FiniteStateMachineEnv(
            num_steps=1, network=ph.Network([mock.MockStrategicAgent("agent")]), initial_stage="A",
            stages=[ph.FSMStage(stage_id="A", acting_agents=["agent"], next_stages=["A"],
                                handler=handler)])

    s = one_stage(ph.StageRule("B")).spec  # "B" is no stage: FSMRuntimeError at run time
    assert s.stages[0].rule_branch[0].then == L.PHX_MAX_STAGES - 1 and s.stages[0].next_allowed == 1
    # an if / elif / else chain: (step >= 2 and step < 5) -> A; elif step == handled-by-column -> A
    rule = ph.StageRule("A", "step", ">=", 2, otherwise="A", also=("step", "<", 5),
                        elifs=[(("step", "==", "step"), "A"), ([("step", "!=", 7), ("step", ">", 0)], "A")])
    g = one_stage(rule).spec.stages[0]
    assert (g.handler, g.rule_n_branches) == (2, 3)
    assert [g.rule_branch[b].n_terms for b in range(3)] == [2, 1, 2]
    assert (g.rule_branch[0].term[1].cmp, g.rule_branch[0].term[1].rhs) == (L.CMP_LT, 5)
    assert g.rule_branch[1].term[0].rhs_kind == L.RULE_STEP
    with pytest.raises(ValueError):
        ph.StageRule("A", also=("step", "<", 5))
    with pytest.raises(ValueError):
        ph.StageRule("A", "step", "<", 1, otherwise="A", elifs=[(("step", "<", k), "A") for k in range(4)])
    with pytest.raises(ph.NotLowerableError):
        one_stage(ph.StageRule("A", "step", "<", ("agent", "nobody", 0), otherwise="A")).spec
    # float32 comparisons: a float32 device column against a float32-exact constant / column
    def echo_stage(handler):
        agents = [mock.MockStrategicAgent("agent"), mock.EchoAgent("e")]
        return ph.FiniteStateMachineEnv(
            num_steps=1, network=ph.Network(agents), initial_stage="A",
            stages=[ph.FSMStage(stage_id="A", acting_agents=["agent"], next_stages=["A"],
                                handler=handler)])

    t = echo_stage(ph.StageRule("A", ("agent", "e", "level"), "<=", 2.25, otherwise="A")
                   ).spec.stages[0].rule_branch[0].term[0]
    assert t.cmp == (L.CMP_LE | L.CMP_F32) and t.rhs_kind == L.RULE_CONST
    assert np.array([t.rhs], np.int32).view(np.float32)[0] == np.float32(2.25)
    t = echo_stage(ph.StageRule("A", ("agent", "e", "level"), ">", ("agent", "e", "level"),
                                otherwise="A")).spec.stages[0].rule_branch[0].term[0]
    assert t.cmp == (L.CMP_GT | L.CMP_F32) and t.rhs_kind == L.RULE_AGENT_WORD and t.rhs_word == 5
    for bad in (ph.StageRule("A", ("agent", "e", "level"), "<", 0.1, otherwise="A"),      # not a float32
                ph.StageRule("A", ("agent", "e", "level"), "<", 3, otherwise="A"),        # float vs int
                ph.StageRule("A", ("agent", "e", "handled_count"), "<", 1.5, otherwise="A"),
                ph.StageRule("A", "step", "<", 1.5, otherwise="A")):
        with pytest.raises(ph.NotLowerableError):
            echo_stage(bad).spec
    with pytest.raises(ph.NotLowerableError):
        one_stage(lambda env: "A").spec
    with pytest.raises(ph.NotLowerableError):
        one_stage(ph.StageRule("A", ("agent", "agent", "no_such_column"), "<", 1, otherwise="A")).spec
    with pytest.raises(ph.NotLowerableError):  # the rule's constant travels as an int32
        one_stage(ph.StageRule("A", "step", "<", 2 ** 31, otherwise="A")).spec
    with pytest.raises(ValueError):
        ph.StageRule("A", "step", "~", 1, otherwise="A")
    with pytest.raises(ph.DeviceOnlyError):
        ph.StageRule("A")()


def test_python_logic_is_never_silently_ignored():
    """An agent class that (re)defines per-message logic in Python cannot be lowered -- the
    product never runs Python handlers on the CPU instead."""
    from phantom_b200.envs.supply_chain import ShopAgent, SupplyChainEnv

    class GreedyShop(ShopAgent):
        def compute_reward(self, ctx):
            return 1.0

    env = SupplyChainEnv()
    env.network.agents["SHOP"].__class__ = GreedyShop
    with pytest.raises(ph.NotLowerableError, match="compute_reward"):
        env.spec

    class Plain(ph.Agent):
        pass

    env = ph.PhantomEnv(num_steps=3, network=ph.Network([Plain("x")]))
    with pytest.raises(ph.NotLowerableError, match="no device program"):
        env.spec
    with pytest.raises(ph.DeviceOnlyError):
        ph.Network([Plain("x")]).send("x", "x", None)
    # shuffle_batches lowers to a flag (resolvers.py:150-151 runs inside the round kernel)
    from phantom_b200 import _lib as L
    from phantom_b200.envs.supply_chain2 import SupplyChain2Env

    spec = SupplyChain2Env(shuffle_batches=True, rates=(0.5, 0.75)).spec
    assert spec.flags & L.FLAG_SHUFFLE_BATCHES and spec.flags & L.FLAG_STOCHASTIC_NETWORK
    assert spec.flags & L.FLAG_IGNORE_CONNECTION_ERRORS
    assert spec.n_base_connections == 2 * 1 + 2 * 5
    assert [spec.base_rate[c] for c in range(12)] == [0.5] * 2 + [0.75] * 10
    assert (spec.base_u[0], spec.base_v[0]) == (1, 0) and (spec.base_u[2], spec.base_v[2]) == (1, 3)


def test_network_construction_api_and_errors():
    """Network builders and their validation errors (network.py:87-177)."""
    from phantom_b200.envs.mock import MockAgent

    with pytest.raises(ValueError):
        ph.Network([MockAgent("a1"), MockAgent("a1")])
    with pytest.raises(ValueError):
        ph.Network([MockAgent("a1")], connections=[("a1", "a2")])
    net = ph.Network([MockAgent("mm"), MockAgent("inv"), MockAgent("inv2")])
    net.add_connection("mm", "inv")
    assert net.has_edge("mm", "inv") and net.has_edge("inv", "mm") and not net.has_edge("mm", "inv2")
    assert len(net) == 3 and net["mm"].id == "mm"
    assert list(net.get_agents_where(lambda a: a.id == "mm")) == ["mm"]
    assert net.get_agents_with_type(ph.Agent) == net.agents and net.get_agents_without_type(ph.Agent) == {}
    ctx = net.context_for("mm", ph.EnvView(0, 0.0))
    assert ctx.neighbour_ids == ["inv"] and "inv" in ctx and "inv2" not in ctx
    sub = net.subnet_for("mm")
    assert list(sub.agents) == ["mm", "inv"]
    with pytest.raises(ValueError, match="square"):
        net.add_connections_with_adjmat(["mm", "inv"], np.zeros((2, 3)))


def test_env_level_hooks_must_be_device_programs():
    """A Python override of an env-level hook cannot run: lowering refuses it (no CPU fallback)."""
    from phantom_b200.envs import simple_market as sm
    from phantom_b200.errors import NotLowerableError

    class Tweaked(sm.SimpleMarketEnv):
        __phx_device_env__ = False

        def post_message_resolution(self):
            self.avg_price = 1.0

    import phantom_b200 as ph
    from phantom_b200.utils.samplers import UniformFloatSampler

    agents = [sm.BuyerAgent("b1", 0.5, supertype=sm.BuyerSupertype(UniformFloatSampler(0.1, 0.2))),
              sm.SellerAgent("s1")]
    net = ph.Network(agents)
    net.add_connections_between(["b1"], ["s1"])
    env = Tweaked(num_steps=4, network=net)
    with pytest.raises(NotLowerableError):
        env.spec  # lowering happens here; no device needed
    assert sm.example_env(num_envs=4).spec.family == L.FAMILY_SIMPLE_MARKET


def test_single_agent_adapter_validation():
    """SingleAgentEnvAdapter's constructor checks (env_wrappers.py:47-68) run before any device
    work, with the reference's messages."""
    from phantom_b200.envs import simple_market as sm

    with pytest.raises(ValueError, match="not found in underlying env"):
        ph.SingleAgentEnvAdapter(sm.example_env, "nobody", {})
    with pytest.raises(ValueError, match="found in agent ID to policy mapping"):
        ph.SingleAgentEnvAdapter(sm.example_env, "s1", {"s1": (ph.Policy, {})})
    with pytest.raises(ValueError, match="has not been defined a policy"):
        ph.SingleAgentEnvAdapter(sm.example_env, "s1", {"b1": (ph.Policy, {})})


def _aligned(shape, dtype, skew=0):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    raw = np.empty(nbytes + 128, np.uint8)
    off = (-raw.ctypes.data) % 64 + skew * np.dtype(dtype).itemsize
    return raw[off:off + nbytes].view(dtype).reshape(shape)


@pytest.mark.parametrize("threads,skew", [(1, 0), (3, 0), (8, 0), (2, 1)])
def test_wire_expansion_equals_the_reference_formulas(threads, skew):
    """Host side of phx_rollout_host's compact supply-chain wire format (no GPU needed): one
    32-bit word per env-step expands to exactly the float32 values the reference computes --
    obs = float32(n / d) (supply_chain.py:124-134), reward = float32(sales - 0.1 * stock)
    (:144-147) -- on the vector path (aligned planes), the scalar path (skewed planes) and the
    ragged tails of every thread's slice."""
    r = np.random.RandomState(threads)
    n = 20011
    stock = r.randint(-32768, 32768, n)
    stock[:4000] = r.randint(0, 101, 4000)
    sales, missed = r.randint(0, 128, n), r.randint(0, 128, n)
    trunc, reset = r.randint(0, 2, n), r.randint(0, 2, n)
    words = (((stock + 32768) & 0xFFFF).astype(np.uint32) | (sales.astype(np.uint32) << 16) |
             (missed.astype(np.uint32) << 23) | (trunc.astype(np.uint32) << 30) |
             (reset.astype(np.uint32) << 31))
    for i in range(0, n, 997):  # the device-side packer agrees on the layout
        assert L.lib.phx_selftest_wire_pack(int(stock[i]), int(sales[i]), int(missed[i]),
                                            int(trunc[i]), int(reset[i])) == words[i]
    obs, rew, ad = _aligned((n, 3), np.float32, skew), _aligned((n,), np.float32, skew), \
        _aligned((n, 2), np.uint8, skew)
    L.check(L.lib.phx_selftest_wire_expand(100, 25, words.ctypes.data, n, threads,
                                           obs.ctypes.data, rew.ctypes.data, ad.ctypes.data))
    shown = np.where(reset == 1, 0, stock)
    want = np.stack([(shown / 100).astype(np.float32), (sales / 25).astype(np.float32),
                     (missed / 25).astype(np.float32)], axis=1)
    assert np.array_equal(obs, want)
    assert np.array_equal(rew, (sales - 0.1 * stock).astype(np.float32))
    assert np.array_equal(ad[:, 1], trunc) and not ad[:, 0].any()


def test_static_schedule_unit_is_generated_and_compiles(tmp_path):
    """phx_selftest_jit_source (no GPU): the specialised unit of the Stackelberg game carries the
    static message schedule -- leaders' turn: one round of 3 Prices; followers' turn: 3 Demands,
    then 3 Acks -- and compiles for sm_100a."""
    import shutil
    import subprocess

    from phantom_b200 import jit
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    env = StackelbergGameEnv(num_envs=256, exec_mode="thread", auto_reset=True)
    spec = env.spec
    need = C.c_uint64(0)
    L.check(L.lib.phx_selftest_jit_source(C.byref(spec), 256, 7, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    L.check(L.lib.phx_selftest_jit_source(C.byref(spec), 256, 7, buf, need.value, None))
    text = buf.value.decode()
    assert "constexpr StaticPlan kPlan" in text and "N_PHASES = 2" in text
    plan = text[text.index("kPlan = {"):]
    rows = [ln.strip() for ln in plan.splitlines()[1:4]]
    assert rows[0] == "2," and rows[1].startswith("{1, 2, 0") and rows[2].startswith("{{3, 0, 0, 0}, {3, 3, 0, 0}")
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    unit = tmp_path / "unit.cu"
    unit.write_text(text)
    proc = subprocess.run([nvcc, *jit.ARCH, "-O3", "-std=c++17", "-cubin", "-I", jit.CSRC, "-o",
                           str(tmp_path / "unit.cubin"), str(unit)], capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]


def test_compound_stage_rule_unit_compiles(tmp_path):
    """The specialised unit of an FSM whose stage handlers are if / elif / else chains
    (StageRule also= / elifs=, include/phx.h phx_rule_branch) is generated without a GPU and
    compiles for sm_100a with the chain folded."""
    import shutil
    import subprocess

    import phantom_b200 as ph
    from phantom_b200 import jit
    from phantom_b200.envs import mock

    class NS:
        pass

    K = NS()
    K.ph, K.MockStrategicAgent, K.EchoAgent = ph, mock.MockStrategicAgent, mock.EchoAgent
    K.finish_network, K.stage_handler = (lambda n: n), ph.StageRule
    from . import kat_scenarios as kats

    env = None
    for seed in range(40):  # the first case with a three-branch chain
        env, _, _ = kats.random_handler_fsm(K, seed, compound=True, num_envs=64, exec_mode="thread")
        if max(st.rule_n_branches for st in env.spec.stages) >= 3:
            break
    spec = env.spec
    need = C.c_uint64(0)
    L.check(L.lib.phx_selftest_jit_source(C.byref(spec), 64, 7, None, 0, C.byref(need)))
    buf = C.create_string_buffer(need.value)
    L.check(L.lib.phx_selftest_jit_source(C.byref(spec), 64, 7, buf, need.value, None))
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    unit = tmp_path / "unit.cu"
    unit.write_text(buf.value.decode())
    proc = subprocess.run([nvcc, *jit.ARCH, "-O3", "-std=c++17", "-cubin", "-Xptxas", "-v", "-I",
                           jit.CSRC, "-o", str(tmp_path / "unit.cubin"), str(unit)],
                          capture_output=True, text=True)
    assert proc.returncode == 0, proc.stderr[-2000:]
    # the chain folds: no local-memory frame beyond the few bytes the run-time acting order of this
    # random env class costs (its stage lists are not ascending in slot)
    frame = int(re.search(r"(\d+) bytes stack frame", proc.stderr).group(1))
    assert frame <= 32, proc.stderr[-1500:]


def test_simple_market_refuses_a_neighbour_order_the_device_would_not_reproduce():
    """SellerAgent.decode_action walks ctx.neighbour_ids in the graph's insertion order
    (market_agents.py:108-112); the device program walks slots.  A network whose connection
    order differs from its agent order is refused at lowering, not mis-ordered silently."""
    import phantom_b200 as ph
    from phantom_b200.envs import simple_market as sm
    from phantom_b200.utils.samplers import UniformFloatSampler

    def build(first, second):
        agents = [sm.BuyerAgent(b, 0.5, supertype=sm.BuyerSupertype(UniformFloatSampler(0.2, 0.2)))
                  for b in ("b1", "b2")] + [sm.SellerAgent("s1")]
        net = ph.Network(agents)
        net.add_connection(first, "s1")
        net.add_connection(second, "s1")
        return sm.SimpleMarketEnv(num_steps=3, network=net)

    assert build("b1", "b2").spec.n_agents == 3
    with pytest.raises(ph.NotLowerableError, match="connected in an order"):
        build("b2", "b1").spec


def test_lowering_keeps_the_acting_order_of_the_users_lists():
    """FSMStage.acting_agents / leader_agents / follower_agents are walked in LIST order by the
    reference (fsm.py:276-277, stackelberg.py:133-140): phx_stage.act_order carries a list that
    is not ascending in slot, an ascending one is left at 0 (= slot order), a duplicate is refused."""
    import phantom_b200 as ph
    from phantom_b200.envs import mock

    def agents():
        return [mock.EchoAgent("e0"), mock.EchoAgent("e1"), mock.MockStrategicAgent("s")]

    def fsm(order):
        return ph.FiniteStateMachineEnv(
            num_steps=2, network=ph.Network(agents()), initial_stage="A",
            stages=[ph.FSMStage(stage_id="A", acting_agents=order, next_stages=["A"])])

    st = fsm(["e1", "s", "e0"]).spec.stages[0]
    assert st.n_act_order == 3 and list(st.act_order[:3]) == [1, 2, 0] and st.acting[0] == 0b111
    assert fsm(["e0", "s"]).spec.stages[0].n_act_order == 0
    with pytest.raises(ph.NotLowerableError, match="listed twice"):
        fsm(["e0", "e0"]).spec
    env = ph.StackelbergEnv(4, ph.Network(agents()), ["s", "e0"], ["e1"])
    assert env.spec.stages[0].n_act_order == 2 and list(env.spec.stages[0].act_order[:2]) == [2, 0]
    assert env.spec.stages[1].n_act_order == 0
