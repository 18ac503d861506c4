"""Observation encoders (reference: phantom/encoders.py:14-131).

In the reference an Encoder is a Python callable `encode(ctx)`.  Here an Encoder *declares* a
small device program: a list of ops `(opcode, length, value)` that the fused kernel evaluates
into the agent's observation row.  Composition (`ChainedEncoder`, `DictEncoder`, `.chain()`)
keeps the reference's semantics -- sub-encoders are evaluated in list / dict order -- and the
tensor API lays the parts out flattened, in that order (the reference returns a tuple / dict
of arrays).  An Encoder subclass without device ops cannot be lowered.
"""
from __future__ import annotations

from collections.abc import Iterable as _Iterable
from typing import Any, Dict, Iterable, List, Mapping, Tuple

import numpy as np

from . import spaces
from .errors import DeviceOnlyError, NotLowerableError

OP_CONST, OP_TIME_ELAPSED, OP_CURRENT_STEP = 0, 1, 2


def flatten(xs: Iterable[Any]) -> List[Any]:
    """utils/__init__.py:14-20"""
    out: List[Any] = []
    for x in xs:
        out.extend(flatten(x) if isinstance(x, _Iterable) else [x])
    return out


class Encoder:
    @property
    def observation_space(self):
        raise NotImplementedError

    def device_ops(self) -> List[Tuple[int, int, float]]:
        raise NotLowerableError(
            f"encoder {type(self).__name__} has no device program (define device_ops())")

    def encode(self, ctx):
        raise DeviceOnlyError("Encoder.encode runs inside the fused step kernel")

    def chain(self, others: Iterable["Encoder"]) -> "ChainedEncoder":
        return ChainedEncoder(flatten([self, others]))

    def reset(self):
        pass

    def flat_dim(self) -> int:
        return sum(n for _, n, _ in self.device_ops())

    def unflatten(self, flat: np.ndarray):
        """Flat device row -> the structure the reference's encode() returns."""
        return np.asarray(flat[: self.flat_dim()], np.float32).reshape(self.observation_space.shape)

    def __repr__(self) -> str:
        return repr(self.observation_space)

    def __str__(self) -> str:
        return str(self.observation_space)


class EmptyEncoder(Encoder):
    """Generates an empty observation: zeros((1,)) (encoders.py:53-61)."""

    @property
    def observation_space(self):
        return spaces.Box(-np.inf, np.inf, (1,))

    def device_ops(self):
        return [(OP_CONST, 1, 0.0)]


class Constant(Encoder):
    """A constant-valued Box (encoders.py:114-131)."""

    def __init__(self, shape: Tuple[int], value: float = 0.0) -> None:
        self._shape, self._value = tuple(shape), float(value)

    @property
    def observation_space(self):
        return spaces.Box(-np.inf, np.inf, shape=self._shape, dtype=np.float32)

    def device_ops(self):
        return [(OP_CONST, int(np.prod(self._shape)), self._value)]


class ElapsedTime(Encoder):
    """[ctx.env_view.proportion_time_elapsed] -- the observation of the reference's own test
    agent (tests/__init__.py:50-52) as a reusable encoder."""

    @property
    def observation_space(self):
        return spaces.Box(0.0, 1.0, (1,))

    def device_ops(self):
        return [(OP_TIME_ELAPSED, 1, 0.0)]


class CurrentStep(Encoder):
    """[float(ctx.env_view.current_step)]"""

    @property
    def observation_space(self):
        return spaces.Box(0.0, np.inf, (1,))

    def device_ops(self):
        return [(OP_CURRENT_STEP, 1, 0.0)]


class ChainedEncoder(Encoder):
    """n encoders -> Tuple space, evaluated in order (encoders.py:64-87)."""

    def __init__(self, encoders: Iterable[Encoder]):
        self.encoders: List[Encoder] = flatten(encoders)

    @property
    def observation_space(self):
        return spaces.Tuple(tuple(e.observation_space for e in self.encoders))

    def device_ops(self):
        return [op for e in self.encoders for op in e.device_ops()]

    def chain(self, others: Iterable[Encoder]) -> "ChainedEncoder":
        return ChainedEncoder(self.encoders + list(others))

    def unflatten(self, flat):
        out, at = [], 0
        for e in self.encoders:
            out.append(e.unflatten(flat[at:]))
            at += e.flat_dim()
        return tuple(out)

    def reset(self):
        for e in self.encoders:
            e.reset()


class DictEncoder(Encoder):
    """name -> encoder, Dict space, evaluated in dict order (encoders.py:90-111)."""

    def __init__(self, encoders: Mapping[str, Encoder]):
        self.encoders: Dict[str, Encoder] = dict(encoders)

    @property
    def observation_space(self):
        return spaces.Dict({k: e.observation_space for k, e in self.encoders.items()})

    def device_ops(self):
        return [op for e in self.encoders.values() for op in e.device_ops()]

    def unflatten(self, flat):
        out, at = {}, 0
        for k, e in self.encoders.items():
            out[k] = e.unflatten(flat[at:])
            at += e.flat_dim()
        return out

    def reset(self):
        for e in self.encoders.values():
            e.reset()
