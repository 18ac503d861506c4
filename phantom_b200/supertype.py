"""Supertype (reference: phantom/supertype.py:14-110): a dataclass whose fields are constants
or Samplers.  On the device every Sampler-valued field becomes a per-env state column that the
reset kernel re-draws each episode; constant fields become family parameters."""
from __future__ import annotations

import dataclasses
from abc import ABC

from .utils.samplers import Sampler


@dataclasses.dataclass
class Supertype(ABC):
    def sample(self) -> "Supertype":
        """Host-side view: Sampler fields resolve to their last known value."""
        out = {}
        for name in self.__dataclass_fields__:
            v = getattr(self, name)
            out[name] = v.value if isinstance(v, Sampler) else v
        return self.__class__(**out)
