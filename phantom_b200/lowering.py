"""Lowering: Python env objects -> flat `phx_spec` (include/phx.h).

Walks exactly the objects the reference's step loop consults:
  agent order            Network.agents insertion order   (phantom/network.py:96, env.py:142)
  strategic agents       isinstance(a, StrategicAgent)    (phantom/env.py:147-159)
  edges                  the directed graph               (phantom/network.py:122-123,224-231)
  payload whitelists     _sender_types / _receiver_types matched against the agent's exact
                         class name                       (phantom/network.py:315-331)
  resolver options       round_limit, enable_tracking     (phantom/resolvers.py:109-120)
  network options        ignore_connection_errors, enforce_msg_payload_checks
  FSM stages             acting / rewarded / next         (phantom/fsm.py:45-63)
  Stackelberg groups     leader_agents / follower_agents  (phantom/stackelberg.py:47-48)
"""
from __future__ import annotations

import ctypes as C
from typing import List

import numpy as np

from . import _lib as L
from . import families
from .agents import Agent, StrategicAgent
from .errors import NotLowerableError
from .resolvers import BatchResolver

_DEVICE_ONLY_METHODS = (
    "handle_batch", "handle_message", "generate_messages", "pre_message_resolution",
    "post_message_resolution", "encode_observation", "decode_action", "compute_reward",
    "is_terminated", "is_truncated",
)


def _check_no_python_logic(agent: Agent) -> None:
    """A user subclass of a device agent class must not (re)define per-message logic in
    Python: it would be silently ignored by the kernel."""
    for klass in type(agent).__mro__:
        if klass.__dict__.get("__phx_device_class__", False):
            return
        if klass in (Agent, StrategicAgent, object):
            break
        for name, attr in klass.__dict__.items():
            if name in _DEVICE_ONLY_METHODS or hasattr(attr, "_message_type"):
                raise NotLowerableError(
                    f"agent class '{klass.__name__}' defines '{name}' in Python but the step "
                    "loop runs on the GPU; only registered device programs can supply "
                    "per-message logic (phantom_b200 has no CPU fallback)")
    raise NotLowerableError(
        f"agent '{agent.id}' of class '{type(agent).__name__}' has no device program "
        "(__phx_family__ / __phx_kind__); phantom_b200 has no CPU fallback")


_ENV_LEVEL_HOOKS = ("pre_message_resolution", "post_message_resolution", "resolve_network", "view")


def _check_env_hooks(env) -> None:
    """An env class may override the env-level hooks of the step loop (phantom/env.py:166-183;
    the reference's simple_market example does).  Such an override runs on the GPU or not at
    all: the class must declare that its device program implements it
    (`__phx_device_env__ = True`); a plain Python override would be silently ignored."""
    from .env import PhantomEnv
    from .fsm import FiniteStateMachineEnv
    from .stackelberg import StackelbergEnv

    for klass in type(env).__mro__:
        if klass in (PhantomEnv, FiniteStateMachineEnv, StackelbergEnv, object):
            break
        if klass.__dict__.get("__phx_device_env__", False):
            return
        for name in _ENV_LEVEL_HOOKS:
            if name in klass.__dict__:
                raise NotLowerableError(
                    f"env class '{klass.__name__}' overrides '{name}' in Python but the step loop "
                    "runs on the GPU; only env classes backed by a device program "
                    "(__phx_device_env__) can supply env-level hooks (no CPU fallback)")


def _lower_stage_rule(out, sid, st, stage_ids, agents, slot) -> None:
    """FSMStage.handler -> phx_stage.rule_* (reference: phantom/fsm.py:294-307)."""
    from .fsm import StageRule

    rule = st.handler
    if not isinstance(rule, StageRule):
        raise NotLowerableError(
            f"FSM stage '{sid}' has a Python handler; only a phantom_b200.fsm.StageRule "
            "(or no handler) lowers to the device")

    def stage_index(x):
        # a stage the FSM does not know can only be *returned* at run time, where the reference
        # raises FSMRuntimeError (fsm.py:304-307): give it an index outside every next_allowed
        return stage_ids.index(x) if x in stage_ids else L.PHX_MAX_STAGES - 1

    returned = [stage for _, stage in rule.branches] + [rule.otherwise]
    for x in returned:
        if x not in stage_ids and len(stage_ids) >= L.PHX_MAX_STAGES:
            raise NotLowerableError("StageRule returns an unknown stage and no spare stage index is left")

    def operand(x, rhs):
        """-> (kind, slot, word, constant, is_float32) of one side of a comparison."""
        if rhs and isinstance(x, (int, np.integer)) and not isinstance(x, bool):
            if not -2 ** 31 <= int(x) < 2 ** 31:
                raise NotLowerableError(f"StageRule of stage '{sid}': rhs {x} is not an int32")
            return L.RULE_CONST, 0, 0, int(x), False
        if rhs and isinstance(x, (float, np.floating)):
            # compared as float32 on the device: the constant must BE a float32, or `<=` / `==`
            # against the Python handler's float64 comparison could differ at that very value
            if not np.isfinite(x) or float(np.float32(x)) != float(x):
                raise NotLowerableError(
                    f"StageRule of stage '{sid}': rhs {x!r} is not exactly representable in float32")
            return L.RULE_CONST, 0, 0, int(np.float32(x).view(np.int32)), True
        if x == "step":
            return L.RULE_STEP, 0, 0, 0, False
        if isinstance(x, tuple) and len(x) == 3 and x[0] == "agent":
            _, aid, column = x
            if aid not in slot:
                raise NotLowerableError(f"StageRule of stage '{sid}': unknown agent '{aid}'")
            is_f32 = False
            if isinstance(column, str):
                desc = getattr(type(agents[slot[aid]]), column, None)
                dname = np.dtype(getattr(desc, "dtype", "int32")).name if hasattr(desc, "word") else ""
                if dname not in ("int32", "float32"):
                    raise NotLowerableError(
                        f"StageRule of stage '{sid}': '{column}' is not an int32 / float32 device "
                        f"column of '{aid}'")
                column, is_f32 = desc.word, dname == "float32"
            return L.RULE_AGENT_WORD, slot[aid], int(column), 0, is_f32
        if isinstance(x, tuple) and len(x) == 2 and x[0] == "env":
            return L.RULE_ENV_WORD, 0, int(x[1]), 0, False
        side = "rhs" if rhs else "lhs"
        raise NotLowerableError(f"StageRule of stage '{sid}': unknown {side} {x!r}")

    out.handler = 2
    out.rule_resolves = int(rule.resolve_network)
    out.rule_else = stage_index(rule.otherwise)
    out.rule_n_branches = len(rule.branches)
    for b, (terms, stage) in enumerate(rule.branches):
        br = out.rule_branch[b]
        br.n_terms, br.then = len(terms), stage_index(stage)
        for k, (lhs, cmp, rhs) in enumerate(terms):
            t = br.term[k]
            t.cmp = StageRule.CMPS.index(cmp)
            if lhs == "always":
                t.lhs = L.RULE_ALWAYS
                continue
            t.lhs, t.slot, t.word, _, lf = operand(lhs, False)
            t.rhs_kind, t.rhs_slot, t.rhs_word, t.rhs, rf = operand(rhs, True)
            if lf != rf:  # an int32 word against a float (or the reverse): bits are not comparable
                raise NotLowerableError(
                    f"StageRule of stage '{sid}': {lhs!r} {cmp} {rhs!r} mixes int32 and float32 operands")
            if lf:
                t.cmp |= L.CMP_F32


def _lower_act_order(stage, ids, slot, what) -> None:
    """The order in which a stage's agents act = the order of the user's list (the reference
    hands it to _handle_acting_agents: fsm.py:276-277, stackelberg.py:133-140).  Stored only when
    it differs from slot order (include/phx.h phx_stage.n_act_order)."""
    slots = [slot[a] for a in ids]
    if len(set(slots)) != len(slots):
        raise NotLowerableError(f"{what}: an agent is listed twice")
    if slots != sorted(slots):
        stage.n_act_order = len(slots)
        for i, v in enumerate(slots):
            stage.act_order[i] = v


def lower(env, exec_mode: str = "auto", auto_reset: bool = False) -> L.PhxSpec:
    from .fsm import FiniteStateMachineEnv
    from .stackelberg import StackelbergEnv

    _check_env_hooks(env)
    net = env.network
    agents: List[Agent] = list(net.agents.values())
    if not agents:
        raise NotLowerableError("env has no agents")
    if len(agents) > L.PHX_MAX_AGENTS:
        raise NotLowerableError(f"more than {L.PHX_MAX_AGENTS} agents per env")
    for a in agents:
        _check_no_python_logic(a)
    fam_names = {type(a).__phx_family__ for a in agents}
    if len(fam_names) != 1 or None in fam_names:
        raise NotLowerableError(f"agents belong to different device families: {fam_names}")
    info = families.get(fam_names.pop())

    spec = L.PhxSpec()
    spec.struct_size = C.sizeof(L.PhxSpec)
    spec.family = info.family_id
    spec.exec_mode = L.EXEC_MODES[exec_mode]
    spec.num_steps = int(env.num_steps)
    spec.n_agents = len(agents)
    spec.obs_dim, spec.act_dim = info.obs_dim, info.act_dim

    slot = {a.id: i for i, a in enumerate(agents)}
    n_strat = 0
    for i, a in enumerate(agents):
        spec.agent_kind[i] = int(type(a).__phx_kind__)
        if isinstance(a, StrategicAgent):
            spec.strategic_index[i] = n_strat
            n_strat += 1
        else:
            spec.strategic_index[i] = -1
        a._phx_slot = i
        for nb in net.neighbours(a.id):
            L.set_mask(spec.adjacency[i], slot[nb])
    spec.n_strategic = n_strat

    # payload whitelists -> per-slot bitmasks
    if len(info.payload_types) > L.PHX_MAX_TYPES:
        raise NotLowerableError("too many payload types")
    spec.n_payload_types = len(info.payload_types)
    for t, cls in enumerate(info.payload_types):
        if not hasattr(cls, "_sender_types") or not hasattr(cls, "_receiver_types"):
            raise NotLowerableError(
                f"payload class {cls.__name__} must use the msg_payload decorator")
        for i, a in enumerate(agents):
            name = type(a).__name__
            if cls._sender_types is None or name in cls._sender_types:
                L.set_mask(spec.type_sender_ok[t], i)
            if cls._receiver_types is None or name in cls._receiver_types:
                L.set_mask(spec.type_receiver_ok[t], i)

    # resolver / network options
    res = net.resolver
    if not isinstance(res, BatchResolver):
        raise NotLowerableError(
            f"resolver {type(res).__name__}: only BatchResolver has a device implementation")
    spec.round_limit = -1 if res.round_limit is None else int(res.round_limit)
    flags = 0
    if net.ignore_connection_errors:
        flags |= L.FLAG_IGNORE_CONNECTION_ERRORS
    if not net.enforce_msg_payload_checks:
        flags |= L.FLAG_NO_PAYLOAD_CHECKS
    if res.enable_tracking:
        flags |= L.FLAG_TRACK_MESSAGES
        spec.trace_capacity = int(info.trace_capacity(env, agents))
    if auto_reset:
        flags |= L.FLAG_AUTO_RESET
    if getattr(res, "shuffle_batches", False):
        flags |= L.FLAG_SHUFFLE_BATCHES
    base = getattr(net, "_base_connections", None)
    if base is not None:  # StochasticNetwork (phantom/network.py:340-453)
        if len(base) > L.PHX_MAX_BASE_CONNECTIONS:
            raise NotLowerableError(
                f"more than {L.PHX_MAX_BASE_CONNECTIONS} StochasticNetwork base connections")
        if len(agents) > L.PHX_MAX_AGENTS:
            raise NotLowerableError(f"StochasticNetwork: at most {L.PHX_MAX_AGENTS} agents per env")
        flags |= L.FLAG_STOCHASTIC_NETWORK
        spec.n_base_connections = len(base)
        for c, (u, v, rate) in enumerate(base):
            spec.base_u[c], spec.base_v[c], spec.base_rate[c] = slot[u], slot[v], float(rate)
    spec.flags = flags

    # step-loop kind
    if isinstance(env, FiniteStateMachineEnv):
        spec.env_kind = L.ENV_FSM
        stage_ids = list(env._stages)
        if len(stage_ids) > L.PHX_MAX_STAGES:
            raise NotLowerableError(f"more than {L.PHX_MAX_STAGES} FSM stages")
        spec.n_stages = len(stage_ids)
        spec.initial_stage = stage_ids.index(env.initial_stage)
        for k, sid in enumerate(stage_ids):
            st = env._stages[sid]
            if st.handler is not None:
                _lower_stage_rule(spec.stages[k], sid, st, stage_ids, agents, slot)
            for aid in st.acting_agents:
                L.set_mask(spec.stages[k].acting, slot[aid])
            _lower_act_order(spec.stages[k], st.acting_agents, slot, f"acting_agents of stage '{sid}'")
            if st.rewarded_agents is None:
                spec.stages[k].rewarded_is_none = 1
            else:
                for aid in st.rewarded_agents:
                    L.set_mask(spec.stages[k].rewarded, slot[aid])
            for nxt in st.next_stages:
                spec.stages[k].next_allowed |= 1 << stage_ids.index(nxt)
            if st.handler is None:
                spec.stages[k].next_stage = stage_ids.index(st.next_stages[0])
    elif isinstance(env, StackelbergEnv):
        spec.env_kind = L.ENV_STACKELBERG
        for aid in env.leader_agents:
            L.set_mask(spec.leaders, slot[aid])
        for aid in env.follower_agents:
            L.set_mask(spec.followers, slot[aid])
        _lower_act_order(spec.stages[0], env.leader_agents, slot, "leader_agents")
        _lower_act_order(spec.stages[1], env.follower_agents, slot, "follower_agents")
    else:
        spec.env_kind = L.ENV_BASE
    if spec.env_kind not in info.env_kinds:
        raise NotLowerableError(
            f"family '{info.name}' has no kernel for env kind {spec.env_kind}")

    uses_supertypes = bool(getattr(env, "_samplers", None)) or any(
        getattr(a, "supertype", None) is not None for a in agents)
    if uses_supertypes and not info.supports_supertypes:
        raise NotLowerableError(
            f"family '{info.name}' has no device program for agent supertypes / samplers")
    info.collect(env, agents, spec)
    return spec
