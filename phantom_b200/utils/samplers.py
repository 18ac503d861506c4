"""Samplers (reference: phantom/utils/samplers.py:48-271).

In the reference a Sampler draws from process-global `np.random` when the env calls
`sampler.sample()` at reset (phantom/env.py:212-216).  Here a Sampler *declares* a distribution;
the reset kernel draws one value PER ENV from the counter-based RNG contract (stream 3, step 0,
idx = the sampler's position in the env's sampler list) and the value lives in a state column.
"""
from __future__ import annotations

from abc import ABC
from typing import Generic, Optional, TypeVar

from ..errors import NotLowerableError

T = TypeVar("T")

KIND_UNIFORM_FLOAT, KIND_UNIFORM_INT = 1, 2


class Sampler(ABC, Generic[T]):
    def __init__(self):
        self._value: Optional[T] = None

    @property
    def value(self) -> Optional[T]:
        """Last sampled value: filled from the device for num_envs == 1, else None."""
        return self._value

    def sample(self) -> T:
        raise NotLowerableError("samplers are drawn per env by the reset kernel")

    def device_desc(self):
        """(kind, low, high)"""
        raise NotLowerableError(f"sampler {type(self).__name__} has no device distribution")


class ComparableSampler(Sampler[T]):
    pass


class UniformFloatSampler(ComparableSampler[float]):
    """low + (high - low) * u, u uniform in [0, 1) (np.random.uniform, samplers.py:119-147)."""

    def __init__(self, low: float = 0.0, high: float = 1.0, clip_low=None, clip_high=None) -> None:
        assert high >= low
        self.low, self.high, self.clip_low, self.clip_high = low, high, clip_low, clip_high
        super().__init__()

    def device_desc(self, allow_clip: bool = False):
        """(kind, low, high); families that implement np.clip on the device (samplers.py:144-145)
        pass allow_clip and read clip_low / clip_high themselves."""
        if not allow_clip and (self.clip_low is not None or self.clip_high is not None):
            raise NotLowerableError("this family does not lower UniformFloatSampler clipping")
        return (KIND_UNIFORM_FLOAT, float(self.low), float(self.high))


class UniformIntSampler(ComparableSampler[int]):
    """np.random.randint(low, high) (samplers.py:150-180)."""

    def __init__(self, low: int = 0, high: int = 1, clip_low=None, clip_high=None) -> None:
        assert high >= low
        self.low, self.high = low, high
        super().__init__()

    def device_desc(self):
        return (KIND_UNIFORM_INT, float(self.low), float(self.high))
