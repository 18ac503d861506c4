"""Agent plugin API (reference: phantom/agents.py:34-349).

In the reference an agent *is* its Python callbacks.  Here an agent class *declares* which
device program implements those callbacks:

    class ShopAgent(ph.StrategicAgent):
        __phx_family__ = "supply_chain"    # device program family (one fused kernel)
        __phx_kind__ = 0                   # agent kind id inside that family

The constructor signatures, `id`, `supertype`, spaces and the overridable method names are
the reference's; per-message methods raise DeviceOnlyError when called on the host, and an
agent class without a device program makes lowering fail (NotLowerableError) instead of
silently running Python.  State columns (e.g. `shop.stock`) are read back from HBM through
`phx_get_field` once the agent is attached to an env.
"""
from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

from .errors import DeviceOnlyError
from .types import AgentID


def msg_handler(message_type):
    """Marks a method as the handler of `message_type` (reference: agents.py:344-349).
    Kept so that agent classes can document their handler table; lowering checks that every
    decorated payload type is one the agent's device program handles."""

    def decorator(fn: Callable) -> Callable:
        fn._message_type = message_type
        return fn

    return decorator


class _DeviceColumn:
    """Descriptor: agent attribute backed by a state word in HBM (int32 or float32)."""

    def __init__(self, word: int, dtype="int32", default=0):
        self.word, self.dtype, self.default = word, dtype, default

    def __set_name__(self, owner, name):
        self.name = name

    def __get__(self, agent, owner=None):
        if agent is None:
            return self
        env = getattr(agent, "_phx_env", None)
        if env is None or not env.is_live:
            return agent.__dict__.get("_col_" + self.name, self.default)
        col = env.agent_column(agent, self.word, self.dtype)
        return col.item() if col.size == 1 else col

    def __set__(self, agent, value):
        env = getattr(agent, "_phx_env", None)
        if env is None or not env.is_live:
            agent.__dict__["_col_" + self.name] = value
        else:
            import numpy as np

            v = np.asarray(value, dtype=self.dtype)
            env.set_agent_column(agent, self.word, v.view(np.int32) if v.dtype != np.int32 else v)


def device_column(word: int, dtype="int32", default=0) -> _DeviceColumn:
    return _DeviceColumn(word, dtype, default)


class Agent:
    __phx_family__: Optional[str] = None
    __phx_kind__: Optional[int] = None

    def __init__(self, agent_id: AgentID, supertype=None) -> None:
        self._id = agent_id
        self.supertype = supertype
        self._phx_env = None
        self._phx_slot = -1

    @property
    def id(self) -> AgentID:
        return self._id

    # -- structural hooks (host side, same meaning as the reference)
    def view(self, neighbour_id: Optional[AgentID] = None):
        return None

    def reset(self) -> None:
        """Host-side per-episode hook; device state is reset by the reset kernel."""

    # -- per-message work: lives in the fused kernel
    def handle_batch(self, ctx, batch):
        raise DeviceOnlyError("Agent.handle_batch runs inside the fused step kernel")

    def handle_message(self, ctx, message):
        raise DeviceOnlyError("Agent.handle_message runs inside the fused step kernel")

    def generate_messages(self, ctx):
        raise DeviceOnlyError("Agent.generate_messages runs inside the fused step kernel")

    def pre_message_resolution(self, ctx) -> None:
        raise DeviceOnlyError("pre_message_resolution runs inside the fused step kernel")

    def post_message_resolution(self, ctx) -> None:
        raise DeviceOnlyError("post_message_resolution runs inside the fused step kernel")

    def __repr__(self) -> str:
        return f"[{self.__class__.__name__} {self.id}]"


class StrategicAgent(Agent):
    def __init__(self, agent_id: AgentID, observation_encoder=None, action_decoder=None,
                 reward_function=None, supertype=None) -> None:
        super().__init__(agent_id, supertype)
        self.observation_encoder = observation_encoder
        self.action_decoder = action_decoder
        self.reward_function = reward_function
        if action_decoder is not None:
            self.action_space = action_decoder.action_space
        elif "action_space" not in dir(self):
            self.action_space = None
        if observation_encoder is not None:
            self.observation_space = observation_encoder.observation_space
        elif "observation_space" not in dir(self):
            self.observation_space = None

    def encode_observation(self, ctx):
        raise DeviceOnlyError("encode_observation runs inside the fused step kernel")

    def decode_action(self, ctx, action):
        raise DeviceOnlyError("decode_action runs inside the fused step kernel")

    def compute_reward(self, ctx) -> float:
        raise DeviceOnlyError("compute_reward runs inside the fused step kernel")

    def is_terminated(self, ctx) -> bool:
        raise DeviceOnlyError("is_terminated runs inside the fused step kernel")

    def is_truncated(self, ctx) -> bool:
        raise DeviceOnlyError("is_truncated runs inside the fused step kernel")

    def collect_infos(self, ctx) -> Dict[str, Any]:
        return {}
