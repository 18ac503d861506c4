"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Encoders, decoders and reward functions of the reference, restated.

Reference lines followed (all under /root/reference/phantom/):
  encoders.py:14-50    Encoder          encoders.py:53-61   EmptyEncoder
  encoders.py:64-87    ChainedEncoder   encoders.py:90-111  DictEncoder
  encoders.py:114-131  Constant
  decoders.py:17-51    Decoder          decoders.py:54-62   EmptyDecoder
  decoders.py:65-93    ChainedDecoder   decoders.py:96-124  DictDecoder
  reward_functions.py:6-38 RewardFunction, Constant
  utils/__init__.py:14-20  flatten
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from collections.abc import Iterable as _Iterable
from typing import Any, Dict, Iterable, List, Mapping, Tuple

import numpy as np

from .core import Context

try:  # gymnasium is absent from the image; the stand-ins are structurally equivalent
    import gymnasium.spaces as _spaces
except ImportError:  # pragma: no cover
    from . import spaces as _spaces


def flatten(xs: Iterable[Any]) -> List[Any]:
    out: List[Any] = []
    for x in xs:
        out.extend(flatten(x) if isinstance(x, _Iterable) else [x])
    return out


# ---------------------------------------------------------------------------- encoders
class Encoder(ABC):
    @property
    @abstractmethod
    def observation_space(self):
        ...

    @abstractmethod
    def encode(self, ctx: Context):
        ...

    def chain(self, others: Iterable["Encoder"]) -> "ChainedEncoder":
        return ChainedEncoder(flatten([self, others]))

    def reset(self):
        pass

    def __repr__(self) -> str:
        return repr(self.observation_space)

    def __str__(self) -> str:
        return str(self.observation_space)


class EmptyEncoder(Encoder):
    @property
    def observation_space(self):
        return _spaces.Box(-np.inf, np.inf, (1,))

    def encode(self, _: Context) -> np.ndarray:
        return np.zeros((1,))


class ChainedEncoder(Encoder):
    def __init__(self, encoders: Iterable[Encoder]):
        self.encoders: List[Encoder] = flatten(encoders)

    @property
    def observation_space(self):
        return _spaces.Tuple(tuple(e.observation_space for e in self.encoders))

    def encode(self, ctx: Context) -> Tuple:
        return tuple(e.encode(ctx) for e in self.encoders)

    def chain(self, others: Iterable[Encoder]) -> "ChainedEncoder":
        return ChainedEncoder(self.encoders + list(others))

    def reset(self):
        for e in self.encoders:
            e.reset()


class DictEncoder(Encoder):
    def __init__(self, encoders: Mapping[str, Encoder]):
        self.encoders: Dict[str, Encoder] = dict(encoders)

    @property
    def observation_space(self):
        return _spaces.Dict({k: e.observation_space for k, e in self.encoders.items()})

    def encode(self, ctx: Context) -> Dict[str, Any]:
        return {k: e.encode(ctx) for k, e in self.encoders.items()}

    def reset(self):
        for e in self.encoders.values():
            e.reset()


class ConstantEncoder(Encoder):
    """encoders.py:114-131 (`Constant`)."""

    def __init__(self, shape: Tuple[int], value: float = 0.0):
        self._shape, self._value = shape, value

    @property
    def observation_space(self):
        return _spaces.Box(-np.inf, np.inf, shape=self._shape, dtype=np.float32)

    def encode(self, _: Context) -> np.ndarray:
        return np.full(self._shape, self._value)


# ---------------------------------------------------------------------------- decoders
class Decoder(ABC):
    @property
    @abstractmethod
    def action_space(self):
        ...

    @abstractmethod
    def decode(self, ctx: Context, action):
        ...

    def chain(self, others: Iterable["Decoder"]) -> "ChainedDecoder":
        return ChainedDecoder(flatten([self, others]))

    def reset(self):
        pass

    def __repr__(self) -> str:
        return repr(self.action_space)

    def __str__(self) -> str:
        return str(self.action_space)


class EmptyDecoder(Decoder):
    @property
    def action_space(self):
        return _spaces.Box(-np.inf, np.inf, (1,))

    def decode(self, _: Context, action):
        return []


class ChainedDecoder(Decoder):
    def __init__(self, decoders: Iterable[Decoder]):
        self.decoders: List[Decoder] = flatten(decoders)

    @property
    def action_space(self):
        return _spaces.Tuple(tuple(d.action_space for d in self.decoders))

    def decode(self, ctx: Context, action: Tuple):
        out = []
        for d, sub in zip(self.decoders, action):
            out.extend(d.decode(ctx, sub))
        return out

    def chain(self, others: Iterable[Decoder]) -> "ChainedDecoder":
        return ChainedDecoder(self.decoders + list(others))

    def reset(self):
        for d in self.decoders:
            d.reset()


class DictDecoder(Decoder):
    def __init__(self, decoders: Mapping[str, Decoder]):
        self.decoders: Dict[str, Decoder] = dict(decoders)

    @property
    def action_space(self):
        return _spaces.Dict({k: d.action_space for k, d in self.decoders.items()})

    def decode(self, ctx: Context, action: Dict[str, Any]):
        out = []
        for k, d in self.decoders.items():
            out.extend(d.decode(ctx, action[k]))
        return out

    def reset(self):
        for d in self.decoders.values():
            d.reset()


# ----------------------------------------------------------------------------- rewards
class RewardFunction(ABC):
    @abstractmethod
    def reward(self, ctx: Context) -> float:
        raise NotImplementedError

    def reset(self):
        pass


class ConstantReward(RewardFunction):
    """reward_functions.py:26-38 (`Constant`)."""

    def __init__(self, value: float = 0.0):
        self.value = value

    def reward(self, _: Context) -> float:
        return self.value
