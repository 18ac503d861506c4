"""PhantomEnv (reference: phantom/env.py:25-351) as a batch of device-resident envs.

Same constructor, properties and `reset()` / `step()` contract as the reference for
`num_envs == 1` (dicts keyed by agent id, `PhantomEnv.Step`), plus the tensor API for
`num_envs >= 1`:

    env = SupplyChainEnv(num_envs=65536, device=0, seed=0)
    obs, obs_mask = env.reset_batch()
    out = env.step_batch(actions)            # BatchStep of torch tensors on the device
    out = env.rollout_batch(actions_T)       # T fused steps in one launch

Every call goes through the C ABI of libphx.so (include/phx.h); torch tensors are only the
buffer currency.  There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Dict, List, Mapping, NamedTuple, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L
from .agents import Agent, StrategicAgent
from .context import Context
from .errors import DeviceOnlyError, NotLowerableError
from .message import Message
from .network import Network, NetworkError
from .types import AgentID
from .views import AgentView, EnvView


class BatchStep(NamedTuple):
    """Tensor form of PhantomEnv.Step.  E envs, S strategic agents (agent order), O obs
    floats; a leading T axis is present for rollouts."""

    observations: Any   # float32 [E,S,O]  valid where obs_mask == 1
    obs_mask: Any       # uint8   [E,S]    agent is a key of Step.observations
    rewards: Any        # float32 [E,S]
    reward_mask: Any    # uint8   [E,S]    0 absent, 1 value, 2 = None
    terminations: Any   # uint8   [E,S]    255 = key absent (agent already done)
    truncations: Any    # uint8   [E,S]
    all_done: Any       # uint8   [E,2]    terminations["__all__"], truncations["__all__"]


def _fault_exception(code: int, env_index: int) -> Exception:
    from .fsm import FSMRuntimeError

    where = f" (env {env_index})"
    return {
        L.FAULT_NO_EDGE: NetworkError("No connection between sender and receiver." + where),
        L.FAULT_BAD_PAYLOAD_TYPE: NetworkError(
            "Message payload type cannot be sent/received by this agent type." + where),
        L.FAULT_UNKNOWN_MSG_TYPE: ValueError("Unknown message type for receiving agent." + where),
        L.FAULT_ROUND_LIMIT: RuntimeError(
            "message(s) still in queue after BatchResolver round limit reached." + where),
        L.FAULT_BAD_TRANSITION: FSMRuntimeError(
            "FiniteStateMachineEnv attempted invalid transition." + where),
        L.FAULT_QUEUE_OVERFLOW: RuntimeError("device message queue capacity exceeded." + where),
        L.FAULT_INVALID_ACTION: ValueError("action is non-finite or outside the contract." + where),
        L.FAULT_UNRESOLVED_MAIL: RuntimeError(
            "an FSM stage handler left messages unresolved; only the thread-per-env engine "
            "(env classes of <= 8 agents) carries mail into a later step." + where),
        L.FAULT_PLAN_MISMATCH: RuntimeError(
            "a statically scheduled step kernel saw a send its device program did not declare "
            "(act_sends / handle_sends); re-create the env without specialise()." + where),
    }.get(code, RuntimeError(f"device fault {code}{where}"))


class PhantomEnv:
    # kernel variant used when an env does not ask for one: "auto" (fastest valid), "fast",
    # "thread" (generic engine, thread per env), "queue" (generic engine, tile per env)
    default_exec_mode = "auto"

    class Step(NamedTuple):
        observations: Dict[AgentID, Any]
        rewards: Dict[AgentID, float]
        terminations: Dict[AgentID, bool]
        truncations: Dict[AgentID, bool]
        infos: Dict[AgentID, Any]

    def __init__(self, num_steps: int, network: Optional[Network] = None,
                 env_supertype=None, agent_supertypes=None, *, num_envs: int = 1,
                 device: int = 0, seed: int = 0, env_offset: int = 0,
                 exec_mode: Optional[str] = None, auto_reset: bool = False) -> None:
        if env_supertype is not None:
            raise NotLowerableError(
                "env supertypes are not lowered to the device yet (SURVEY.md 8f row 2)")
        self.network = network or Network()
        # env.py:78-124: collect every Sampler of the agent supertypes once, in dict order
        self._samplers: List[Any] = []
        if agent_supertypes is not None:
            from .utils.samplers import Sampler

            for agent_id, st in agent_supertypes.items():
                agent = self.network.agents[agent_id]
                if isinstance(st, dict):
                    st = agent.Supertype(**st)
                st._managed = True
                for value in st.__dict__.values():
                    if isinstance(value, Sampler) and not any(value is x for x in self._samplers):
                        self._samplers.append(value)
                agent.supertype = st
        self.num_steps = num_steps
        self.env_supertype = None
        self.env_type = None
        self.num_envs = int(num_envs)
        self.device = int(device)
        self.seed = int(seed)
        self.env_offset = int(env_offset)
        self.exec_mode = exec_mode or PhantomEnv.default_exec_mode
        self.auto_reset = bool(auto_reset)
        self._handle: Optional[C.c_void_p] = None
        self._spec: Optional[L.PhxSpec] = None
        self._out: Optional[BatchStep] = None
        self._step_cache_E1 = 0

    # ------------------------------------------------------------ reference properties
    @property
    def current_step(self):
        if not self.is_live:
            return 0
        col = self.field(L.FIELD_STEP, np.int32)
        return int(col[0]) if self.num_envs == 1 else col

    @property
    def n_agents(self) -> int:
        return len(self.agent_ids)

    @property
    def agents(self) -> Dict[AgentID, Agent]:
        return self.network.agents

    @property
    def agent_ids(self) -> List[AgentID]:
        return list(self.network.agent_ids)

    @property
    def strategic_agents(self) -> List[StrategicAgent]:
        return [a for a in self.agents.values() if isinstance(a, StrategicAgent)]

    @property
    def non_strategic_agents(self) -> List[Agent]:
        return [a for a in self.agents.values() if not isinstance(a, StrategicAgent)]

    @property
    def strategic_agent_ids(self) -> List[AgentID]:
        return [a.id for a in self.strategic_agents]

    @property
    def non_strategic_agent_ids(self) -> List[AgentID]:
        return [a.id for a in self.non_strategic_agents]

    def __getitem__(self, agent_id: AgentID) -> Agent:
        return self.network[agent_id]

    def view(self, agent_views: Dict[AgentID, AgentView]) -> EnvView:
        step = self.current_step if self.num_envs == 1 else 0
        return EnvView(step, step / self.num_steps)

    def render(self) -> None:
        return None

    # env-level hooks of the step loop (phantom/env.py:170-183).  They run inside the fused
    # kernel; an env class that overrides one must be backed by a device program that
    # implements the override (class attribute __phx_device_env__ = True), see lowering.py.
    def pre_message_resolution(self) -> None:
        raise DeviceOnlyError("pre_message_resolution runs inside the fused step kernel")

    def post_message_resolution(self) -> None:
        raise DeviceOnlyError("post_message_resolution runs inside the fused step kernel")

    def resolve_network(self) -> None:
        raise DeviceOnlyError("resolve_network runs inside the fused step kernel")

    # ---------------------------------------------------------------- device plumbing
    @property
    def is_live(self) -> bool:
        return self._handle is not None

    @property
    def spec(self) -> L.PhxSpec:
        if self._spec is None:
            from .lowering import lower

            self._spec = lower(self, self.exec_mode, self.auto_reset)
        return self._spec

    @property
    def exec_name(self) -> str:
        self._ensure_handle()
        return L.lib.phx_exec_name(self._handle).decode()

    def _torch(self):
        import torch

        if not torch.cuda.is_available():
            raise RuntimeError(
                "phantom_b200 needs a CUDA device: the env step loop only exists as CUDA "
                "kernels (no CPU fallback)")
        return torch

    def _ensure_handle(self):
        if self._handle is not None:
            return
        spec = self.spec
        h = C.c_void_p()
        info = self.family
        if info.program_source is not None:  # the user's own device program (csrc/phx_user.cuh)
            from . import jit

            cubin = jit.compile_program(info.program_source)
            L.check(L.lib.phx_create_user(C.byref(spec), cubin.encode(), self.num_envs,
                                          self.device, self.seed, self.env_offset, C.byref(h)))
        else:
            L.check(L.lib.phx_create(C.byref(spec), self.num_envs, self.device, self.seed,
                                     self.env_offset, C.byref(h)))
        self._handle = h
        for a in self.agents.values():
            a._phx_env = self

    def specialise(self) -> "PhantomEnv":
        """Switch this handle to a build of the step kernel specialised to its env class
        (phantom_b200/jit.py; thread-per-env engine).  Same results, fewer instructions."""
        from . import jit

        jit.specialise(self)
        return self

    def close(self) -> None:
        if self._handle is not None:
            L.lib.phx_destroy(self._handle)
            self._handle = None
            self._out = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        torch = self._torch()
        return torch.cuda.current_stream(self.device).cuda_stream

    def _alloc_outputs(self, lead: Tuple[int, ...]) -> BatchStep:
        torch = self._torch()
        E, S, O = self.num_envs, max(self.spec.n_strategic, 1), self.spec.obs_dim
        dev = torch.device("cuda", self.device)
        u8 = lambda *s: torch.zeros(lead + s, dtype=torch.uint8, device=dev)
        return BatchStep(
            torch.zeros(lead + (E, S, O), dtype=torch.float32, device=dev), u8(E, S),
            torch.zeros(lead + (E, S), dtype=torch.float32, device=dev), u8(E, S),
            u8(E, S), u8(E, S), u8(E, 2))

    @staticmethod
    def _ptr(t) -> Optional[int]:
        return None if t is None else t.data_ptr()

    # ---------------------------------------------------------------------- batch API
    def reset_batch(self, env_mask=None):
        """phx_reset: returns (obs [E,S,O] float32, obs_mask [E,S] uint8) device tensors."""
        torch = self._torch()
        self._ensure_handle()
        if self._out is None:
            self._out = self._alloc_outputs(())
        if env_mask is not None:  # (a tensor or anything array-like)
            env_mask = torch.as_tensor(env_mask).to(device=self._out.observations.device,
                                                    dtype=torch.uint8).contiguous()
            if env_mask.numel() != self.num_envs:
                raise ValueError(f"env_mask must have {self.num_envs} entries")
        self._out.obs_mask.zero_()
        L.check(L.lib.phx_reset(self._handle, self._ptr(env_mask),
                                self._out.observations.data_ptr(),
                                self._out.obs_mask.data_ptr(), self._stream()))
        return self._out.observations, self._out.obs_mask

    def _prep_actions(self, actions, action_mask, lead: Tuple[int, ...]):
        torch = self._torch()
        dev = torch.device("cuda", self.device)
        E, S, A = self.num_envs, max(self.spec.n_strategic, 1), self.spec.act_dim
        actions = torch.as_tensor(actions, dtype=torch.float32, device=dev).contiguous()
        if actions.numel() != int(np.prod(lead)) * E * S * A:
            raise ValueError(f"actions must have {lead + (E, S, A)} elements, got {tuple(actions.shape)}")
        if action_mask is not None:
            action_mask = torch.as_tensor(action_mask, device=dev).to(torch.uint8).contiguous()
            if action_mask.numel() != int(np.prod(lead)) * E * S:
                raise ValueError("action_mask must have shape [.., E, S]")
        return actions, action_mask

    def step_batch(self, actions, action_mask=None) -> BatchStep:
        """phx_step.  The returned tensors are the env's persistent output buffers: they are
        overwritten by the next call."""
        self._ensure_handle()
        if self._out is None:
            self._out = self._alloc_outputs(())
        actions, action_mask = self._prep_actions(actions, action_mask, ())
        o = self._out
        L.check(L.lib.phx_step(self._handle, actions.data_ptr(), self._ptr(action_mask),
                               o.observations.data_ptr(), o.obs_mask.data_ptr(),
                               o.rewards.data_ptr(), o.reward_mask.data_ptr(),
                               o.terminations.data_ptr(), o.truncations.data_ptr(),
                               o.all_done.data_ptr(), self._stream()))
        return o

    def rollout_batch(self, actions, action_mask=None, out: Optional[BatchStep] = None) -> BatchStep:
        """phx_rollout: actions [T,E,S,A] -> BatchStep with a leading T axis."""
        self._ensure_handle()
        T = int(actions.shape[0])
        actions, action_mask = self._prep_actions(actions, action_mask, (T,))
        o = out if out is not None else self._alloc_outputs((T,))
        L.check(L.lib.phx_rollout(self._handle, T, actions.data_ptr(), self._ptr(action_mask),
                                  o.observations.data_ptr(), o.obs_mask.data_ptr(),
                                  o.rewards.data_ptr(), o.reward_mask.data_ptr(),
                                  o.terminations.data_ptr(), o.truncations.data_ptr(),
                                  o.all_done.data_ptr(), self._stream()))
        return o

    def rollout_host(self, actions: np.ndarray, action_mask: Optional[np.ndarray] = None,
                     out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """phx_rollout_host: numpy in, numpy out, copies inside the call (the e2e path)."""
        self._ensure_handle()
        T = int(actions.shape[0])
        E, S, O = self.num_envs, max(self.spec.n_strategic, 1), self.spec.obs_dim
        actions = np.ascontiguousarray(actions, np.float32)
        if out is None:
            out = {
                "observations": np.empty((T, E, S, O), np.float32),
                "obs_mask": np.empty((T, E, S), np.uint8),
                "rewards": np.empty((T, E, S), np.float32),
                "reward_mask": np.empty((T, E, S), np.uint8),
                "terminations": np.empty((T, E, S), np.uint8),
                "truncations": np.empty((T, E, S), np.uint8),
                "all_done": np.empty((T, E, 2), np.uint8),
            }
        am = None if action_mask is None else np.ascontiguousarray(action_mask, np.uint8)
        p = lambda a: None if a is None else a.ctypes.data
        L.check(L.lib.phx_rollout_host(
            self._handle, T, p(actions), p(am), p(out["observations"]), p(out["obs_mask"]),
            p(out["rewards"]), p(out["reward_mask"]), p(out["terminations"]),
            p(out["truncations"]), p(out["all_done"])))
        return out

    def check_errors(self, clear: bool = True) -> None:
        """phx_poll_errors: raise the reference exception matching the first device fault."""
        if self._handle is None:
            return
        n, first, code = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(L.lib.phx_poll_errors(self._handle, C.byref(n), C.byref(first), C.byref(code),
                                      1 if clear else 0))
        if n.value:
            raise _fault_exception(code.value, first.value)

    def field(self, field: int, dtype, index: int = 0, width: int = 1) -> np.ndarray:
        """phx_get_field -> numpy array [E] (or [E, width])."""
        self._ensure_handle()
        shape = (self.num_envs,) if width == 1 else (self.num_envs, width)
        buf = np.empty(shape, dtype)
        L.check(L.lib.phx_get_field(self._handle, field, index, buf.ctypes.data, buf.nbytes))
        return buf

    def set_field(self, field: int, values: np.ndarray, index: int = 0) -> None:
        self._ensure_handle()
        values = np.ascontiguousarray(values)
        L.check(L.lib.phx_set_field(self._handle, field, index, values.ctypes.data, values.nbytes))

    # agent attribute <-> state column plumbing (used by agents.device_column)
    @property
    def family(self):
        from . import families

        return families.get(type(next(iter(self.agents.values()))).__phx_family__)

    @property
    def tile_width(self) -> int:
        """Lanes per env of the queue engine (state columns are [E, G], slot-major)."""
        import re

        m = re.search(r"G=(\d+)", self.exec_name)
        return int(m.group(1)) if m else 0

    def adjacency(self) -> np.ndarray:
        """Per-env graphs of a StochasticNetwork: uint8 [E, n_agents, n_agents], entry [e, s, r]
        = 1 iff env e currently has the edge s -> r (network.py:439-448 run per env)."""
        n = self.spec.n_agents
        if self.exec_name.startswith("wide"):  # block engine: rows of PHX_MASK_WORDS words
            W = L.PHX_MASK_WORDS
            rows = self.field(L.FIELD_ADJACENCY, np.uint32, width=self.tile_width * W)
            rows = rows.reshape(rows.shape[0], self.tile_width, W)[:, :n]
            r = np.arange(n)
            return ((rows[:, :, r >> 5] >> (r & 31).astype(np.uint32)[None, None, :]) & 1).astype(np.uint8)
        rows = self.field(L.FIELD_ADJACENCY, np.uint32, width=self.tile_width)
        return ((rows[:, :n, None] >> np.arange(n, dtype=np.uint32)[None, None, :]) & 1).astype(np.uint8)

    def agent_column(self, agent: Agent, word: int, dtype=np.int32) -> np.ndarray:
        """State word `word` of `agent` for every env: array [E]."""
        info = self.family
        if info.fast_column is not None and self.exec_name.startswith("fast"):
            return info.fast_column(self, agent, word)
        col = self.field(L.FIELD_FAMILY + word, np.int32, width=self.tile_width)
        return col[:, agent._phx_slot].view(dtype)

    def reduce_agent_column(self, agent: Agent, word: int) -> Dict[str, float]:
        """phx_reduce_field: {sum, min, max, mean} of an agent's int32 state word over all envs,
        computed on the device (no [E] column crosses PCIe)."""
        self._ensure_handle()
        info = self.family
        if info.fast_column is not None and self.exec_name.startswith("fast"):
            field, width, col = L.FIELD_FAMILY + 0, 4, word
        else:
            field, width, col = L.FIELD_FAMILY + word, self.tile_width, agent._phx_slot
        s, lo, hi = C.c_int64(), C.c_int32(), C.c_int32()
        L.check(L.lib.phx_reduce_field(self._handle, field, 0, width, col, C.byref(s),
                                       C.byref(lo), C.byref(hi)))
        return {"sum": float(s.value), "min": float(lo.value), "max": float(hi.value),
                "mean": s.value / self.num_envs}

    def set_agent_column(self, agent: Agent, word: int, value) -> None:
        info = self.family
        if info.fast_column is not None and self.exec_name.startswith("fast"):
            info.fast_column(self, agent, word, value)
            return
        col = self.field(L.FIELD_FAMILY + word, np.int32, width=self.tile_width)
        col[:, agent._phx_slot] = np.asarray(value).astype(col.dtype)
        self.set_field(L.FIELD_FAMILY + word, col)

    def tracked_messages_batch(self, env_begin: int = 0, env_end: Optional[int] = None,
                               step: Optional[int] = None):
        """phx_get_trace / phx_get_trace_step: (counts [n], rows [n, cap, 4]) of step `step` of
        the last tracked launch (default: its last step)."""
        self._ensure_handle()
        env_end = self.num_envs if env_end is None else env_end
        n, cap = env_end - env_begin, self.spec.trace_capacity
        counts = np.zeros(n, np.int32)
        rows = np.zeros((n, cap, L.PHX_TRACE_WORDS), np.int32)
        if step is None:
            step = L.lib.phx_trace_steps(self._handle) - 1
        L.check(L.lib.phx_get_trace_step(self._handle, int(step), env_begin, env_end,
                                         counts.ctypes.data, rows.ctypes.data))
        return counts, rows

    def rollout_messages(self, env_index: int = 0) -> List[List[Message]]:
        """Resolver.tracked_messages of every step of the last tracked rollout, for one env:
        [step][message] in the reference's push order (phantom/resolvers.py:41-60)."""
        self._ensure_handle()
        return [self._decode_trace(t, env_index) for t in range(L.lib.phx_trace_steps(self._handle))]

    # --------------------------------------------------- reference API (num_envs == 1)
    def _require_single(self, what: str) -> None:
        if self.num_envs != 1:
            raise TypeError(
                f"{what}() with per-agent dicts is the num_envs == 1 view; use {what}_batch() "
                f"for num_envs = {self.num_envs}")

    def _obs_dict(self, obs, obs_mask) -> Dict[AgentID, Any]:
        out = {}
        for s, agent in enumerate(self.strategic_agents):
            if obs_mask[s]:
                enc = getattr(agent, "observation_encoder", None)
                if enc is not None:  # Chained / Dict encoders: rebuild the tuple / dict
                    out[agent.id] = enc.unflatten(np.array(obs[s], dtype=np.float32))
                else:
                    n = self._agent_obs_dim(agent)
                    out[agent.id] = np.array(obs[s, :n], dtype=np.float32)
        return out

    def _agent_obs_dim(self, agent) -> int:
        space = getattr(agent, "observation_space", None)
        shape = getattr(space, "shape", None)
        return int(np.prod(shape)) if shape else self.spec.obs_dim

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        self._require_single("reset")
        if seed is not None and int(seed) != self.seed:
            self.close()
            self.seed = int(seed)
        obs, mask = self.reset_batch()
        self.network.reset()
        return self._obs_dict(obs[0].cpu().numpy(), mask[0].cpu().numpy()), {}

    def _actions_from_mapping(self, actions: Mapping[AgentID, Any]):
        S, A = max(self.spec.n_strategic, 1), self.spec.act_dim
        a = np.zeros((1, S, A), np.float32)
        m = np.zeros((1, S), np.uint8)
        for s, agent in enumerate(self.strategic_agents):
            if agent.id in actions:
                v = np.asarray(actions[agent.id], np.float32).reshape(-1)
                a[0, s, : v.size] = v[:A]
                m[0, s] = 1
        return a, m

    def step(self, actions: Mapping[AgentID, Any]) -> "PhantomEnv.Step":
        self._require_single("step")
        a, m = self._actions_from_mapping(actions)
        out = self.step_batch(a, m)
        host = [t[0].cpu().numpy() for t in out]
        self.check_errors()
        if self.network.resolver.enable_tracking:
            self.network.resolver._tracked_messages.extend(self._decode_trace())
        return self._step_from_host(*host)

    def _step_from_host(self, obs, obs_mask, rew, rew_mask, term, trunc, all_done):
        observations = self._obs_dict(obs, obs_mask)
        rewards, terminations, truncations = {}, {}, {}
        for s, agent in enumerate(self.strategic_agents):
            if rew_mask[s] == 1:
                rewards[agent.id] = float(rew[s])
            elif rew_mask[s] == 2:
                rewards[agent.id] = None
            if term[s] != 255:
                terminations[agent.id] = bool(term[s])
                truncations[agent.id] = bool(trunc[s])
        infos = {aid: {} for aid in observations}
        terminations["__all__"] = bool(all_done[0])
        truncations["__all__"] = bool(all_done[1])
        return self.Step(observations, rewards, terminations, truncations, infos)

    def _decode_trace(self, step: Optional[int] = None, env_index: int = 0) -> List[Message]:
        info = self.family
        counts, rows = self.tracked_messages_batch(env_index, env_index + 1, step)
        ids = self.agent_ids
        msgs = []
        for r in rows[0, : counts[0]]:
            sender, recv, ptype = r[0] & 0xFF, (r[0] >> 8) & 0xFF, (r[0] >> 16) & 0xFF
            cls = info.payload_types[ptype]
            names = [f.name for f in __import__("dataclasses").fields(cls)]
            if hasattr(cls, "_phx_decode"):  # payloads that are not plain int fields
                payload = cls._phx_decode(int(r[1]), int(r[2]))
            else:
                payload = cls(*[int(r[1]), int(r[2])][: len(names)])
            msgs.append(Message(ids[sender], ids[recv], payload))
        return msgs

    def is_terminated(self) -> bool:
        self._require_single("is_terminated")
        term = self.field(L.FIELD_TERMINATED, np.uint32, width=self._done_words)
        return sum(bin(int(w)).count("1") for w in np.atleast_1d(term[0])) == len(self.strategic_agents)

    def is_truncated(self) -> bool:
        self._require_single("is_truncated")
        trunc = self.field(L.FIELD_TRUNCATED, np.uint32, width=self._done_words)
        at_max = self.num_steps is not None and self.current_step == self.num_steps
        return at_max or sum(bin(int(w)).count("1") for w in np.atleast_1d(trunc[0])) == len(self.strategic_agents)

    @property
    def _done_words(self) -> int:
        """Words per env of the done sets: one (bitmask over <= 32 slots), PHX_MASK_WORDS on the
        128-lane block engine."""
        return L.PHX_MASK_WORDS if self.exec_name.startswith("wide") else 1
