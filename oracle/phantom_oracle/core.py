"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Value types and the agent plugin API of the reference, restated.

Reference lines followed (all under /root/reference/phantom/):
  message.py:10-53   MsgPayload, msg_payload, Message
  views.py:5-34      View / AgentView / EnvView
  context.py:11-40   Context
  agents.py:34-175   Agent (handler registry, handle_batch, handle_message, reset)
  agents.py:181-338  StrategicAgent
  agents.py:344-349  msg_handler
  supertype.py:14-30 Supertype.sample
  utils/samplers.py:48-80 Sampler base
"""
from __future__ import annotations

import dataclasses

import numpy as np
from abc import ABC, abstractmethod
from typing import Any, Dict, Generic, Hashable, List, Optional, Sequence, Tuple, TypeVar

AgentID = Hashable
StageID = Hashable
PolicyID = Hashable

T = TypeVar("T")


# ------------------------------------------------------------------ payloads/messages
@dataclasses.dataclass(frozen=True)
class MsgPayload:
    """Deprecated payload base class (message.py:10-12)."""


def _type_names(arg) -> Optional[List[str]]:
    # message.py:22-35 -- None stays None; scalars are wrapped; classes become names.
    if arg is None:
        return None
    items = arg if isinstance(arg, list) else [arg]
    return [t.__name__ if isinstance(t, type) else t for t in items]


def msg_payload(sender_type=None, receiver_type=None):
    """Class decorator: frozen dataclass + sender/receiver class-name whitelists
    (message.py:20-42)."""

    def wrap(cls):
        cls._sender_types = _type_names(sender_type)
        cls._receiver_types = _type_names(receiver_type)
        return dataclasses.dataclass(frozen=True)(cls)

    return wrap


@dataclasses.dataclass(frozen=True)
class Message(Generic[T]):
    """message.py:45-53"""

    sender_id: AgentID
    receiver_id: AgentID
    payload: Any


# ------------------------------------------------------------------------------ views
@dataclasses.dataclass(frozen=True)
class View(ABC):
    """views.py:5-17"""


@dataclasses.dataclass(frozen=True)
class AgentView(View):
    """views.py:20-24"""


@dataclasses.dataclass(frozen=True)
class EnvView(View):
    """views.py:27-34"""

    current_step: int
    proportion_time_elapsed: float


@dataclasses.dataclass(frozen=True)
class Context:
    """context.py:11-40: focal agent (live object), neighbour views (snapshot), env view."""

    agent: "Agent"
    agent_views: Dict[AgentID, Optional[AgentView]]
    env_view: EnvView

    @property
    def neighbour_ids(self) -> List[AgentID]:
        return list(self.agent_views.keys())

    def __getitem__(self, view_id):
        return self.agent_views[view_id]

    def __contains__(self, view_id) -> bool:
        return view_id in self.agent_views


# --------------------------------------------------------------------------- samplers
class Sampler(ABC, Generic[T]):
    """utils/samplers.py:48-80 (only what the env loop touches: value / sample)."""

    def __init__(self):
        self._value: Optional[T] = None

    @property
    def value(self) -> Optional[T]:
        return self._value

    @abstractmethod
    def sample(self) -> T:
        raise NotImplementedError


class ComparableSampler(Sampler[T]):
    """utils/samplers.py:83-116: comparisons act on the last sampled value."""

    def __lt__(self, other):
        return self.value < other

    def __le__(self, other):
        return self.value <= other

    def __gt__(self, other):
        return self.value > other

    def __ge__(self, other):
        return self.value >= other

    def __eq__(self, other):
        if isinstance(other, ComparableSampler):
            return object.__eq__(self, other)
        return self.value == other

    def __ne__(self, other):
        if isinstance(other, ComparableSampler):
            return object.__ne__(self, other)
        return self.value != other

    __hash__ = object.__hash__


class UniformFloatSampler(ComparableSampler[float]):
    """utils/samplers.py:119-147: np.random.uniform(low, high) on every sample()."""

    def __init__(self, low: float = 0.0, high: float = 1.0, clip_low=None, clip_high=None):
        assert high >= low
        self.low, self.high, self.clip_low, self.clip_high = low, high, clip_low, clip_high
        super().__init__()

    def sample(self) -> float:
        import numpy as np

        self._value = np.random.uniform(self.low, self.high)
        if self.clip_low is not None or self.clip_high is not None:
            self._value = np.clip(self._value, self.clip_low, self.clip_high)
        return self._value


@dataclasses.dataclass
class Supertype(ABC):
    """supertype.py:14-30: sample() resolves Sampler-valued fields; env-managed
    supertypes read the sampler's current value instead of re-sampling."""

    def sample(self) -> "Supertype":
        out = {}
        for name in self.__dataclass_fields__:
            v = getattr(self, name)
            if isinstance(v, Sampler):
                v = v.value if hasattr(self, "_managed") else v.sample()
            out[name] = v
        return self.__class__(**out)

    def to_obs_space_compatible_type(self):
        """supertype.py:32-41,64-85: fields as observation-space compatible values (numbers
        become float32 arrays of shape (1,))."""

        def conv(name, obj):
            if isinstance(obj, dict):
                return {k: conv(k, v) for k, v in obj.items()}
            if isinstance(obj, (float, int)):
                return np.array([obj], dtype=np.float32)
            if isinstance(obj, list):
                return [conv(f"{name}[{i}]", v) for i, v in enumerate(obj)]
            if isinstance(obj, tuple):
                return tuple(conv(f"{name}[{i}]", v) for i, v in enumerate(obj))
            if isinstance(obj, np.ndarray):
                return obj
            raise ValueError(
                f"Can't encode field '{name}' with type '{type(obj)}' into obs space compatible type")

        return {name: conv(name, getattr(self, name)) for name in self.__dataclass_fields__}


# ----------------------------------------------------------------------------- agents
def msg_handler(message_type):
    """agents.py:344-349"""

    def decorator(fn):
        fn._message_type = message_type
        return fn

    return decorator


class Agent(ABC):
    def __init__(self, agent_id: AgentID, supertype: Optional[Supertype] = None):
        self._id = agent_id
        self.supertype = supertype
        # agents.py:69-79 -- handlers are discovered by scanning dir(self), i.e. in
        # ALPHABETICAL attribute-name order, skipping the two space attributes.
        self._handlers: Dict[type, list] = {}
        for name in dir(self):
            if name in ("observation_space", "action_space"):
                continue
            attr = getattr(self, name)
            if callable(attr) and hasattr(attr, "_message_type"):
                self._handlers.setdefault(attr._message_type, []).append(attr)

    @property
    def id(self) -> AgentID:
        return self._id

    def view(self, neighbour_id: Optional[AgentID] = None) -> Optional[AgentView]:
        return None  # agents.py:86-88

    def pre_message_resolution(self, ctx: Context) -> None:
        pass

    def post_message_resolution(self, ctx: Context) -> None:
        pass

    def handle_batch(self, ctx: Context, batch: Sequence[Message]):
        # agents.py:110-120: sequential, responses concatenated in message order.
        out = []
        for message in batch:
            responses = self.handle_message(ctx, message)
            if responses is not None:
                out += responses
        return out

    def handle_message(self, ctx: Context, message: Message):
        # agents.py:138-155: exact payload-type lookup (no subclass match), every
        # registered handler runs, None results are dropped, the rest chained.
        ptype = type(message.payload)
        if ptype not in self._handlers:
            raise ValueError(
                f"Unknown message type {ptype} in message sent from "
                f"'{message.sender_id}' to '{self.id}'."
            )
        out = []
        for handler in self._handlers[ptype]:
            r = handler(ctx, message)
            if r is not None:
                out.extend(r)
        return out

    def generate_messages(self, ctx: Context):
        return []  # agents.py:157-158

    def reset(self) -> None:
        # agents.py:160-175
        if self.supertype is not None:
            self.type = self.supertype.sample()
        elif hasattr(self, "Supertype"):
            try:
                self.type = self.Supertype().sample()
            except TypeError as e:
                raise Exception(
                    f"Tried to initialise agent {self.id}'s Supertype with default "
                    f"values but failed:\n\t{e}"
                )

    def __repr__(self) -> str:
        return f"[{self.__class__.__name__} {self.id}]"


class StrategicAgent(Agent):
    def __init__(
        self,
        agent_id: AgentID,
        observation_encoder=None,
        action_decoder=None,
        reward_function=None,
        supertype: Optional[Supertype] = None,
    ):
        super().__init__(agent_id, supertype)
        self.observation_encoder = observation_encoder
        self.action_decoder = action_decoder
        self.reward_function = reward_function
        # agents.py:212-220
        if action_decoder is not None:
            self.action_space = action_decoder.action_space
        elif "action_space" not in dir(self):
            self.action_space = None
        if observation_encoder is not None:
            self.observation_space = observation_encoder.observation_space
        elif "observation_space" not in dir(self):
            self.observation_space = None

    def encode_observation(self, ctx: Context):
        if self.observation_encoder is None:  # agents.py:240-243
            raise NotImplementedError(
                f"Agent '{self.id}' has no Encoder and no encode_observation override"
            )
        return self.observation_encoder.encode(ctx)

    def decode_action(self, ctx: Context, action):
        if self.action_decoder is None:  # agents.py:265-268
            raise NotImplementedError(
                f"Agent '{self.id}' has no Decoder and no decode_action override"
            )
        return self.action_decoder.decode(ctx, action)

    def compute_reward(self, ctx: Context) -> float:
        if self.reward_function is None:  # agents.py:285-288
            raise NotImplementedError(
                f"Agent '{self.id}' has no RewardFunction and no compute_reward override"
            )
        return self.reward_function.reward(ctx)

    def is_terminated(self, ctx: Context) -> bool:
        return False  # agents.py:307

    def is_truncated(self, ctx: Context) -> bool:
        return False  # agents.py:323

    def collect_infos(self, ctx: Context) -> Dict[str, Any]:
        return {}  # agents.py:338


MessageList = List[Tuple[AgentID, Any]]
