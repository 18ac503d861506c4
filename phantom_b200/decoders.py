"""Action decoders (reference: phantom/decoders.py:17-124).

A Decoder declares the device program that turns an agent's action row into messages.  The
built-in device decoder is `EmptyDecoder` (no messages); workload families supply their own
(e.g. the shop's restock request is part of the supply-chain device program).  Composition
keeps the reference's order: `ChainedDecoder` consumes consecutive action slices in list order,
`DictDecoder` in dict order.
"""
from __future__ import annotations

from typing import Any, Dict, Iterable, List, Mapping

import numpy as np

from . import spaces
from .encoders import flatten
from .errors import DeviceOnlyError, NotLowerableError

OP_NO_MESSAGES = 0


class Decoder:
    @property
    def action_space(self):
        raise NotImplementedError

    def device_ops(self) -> List[tuple]:
        """[(opcode, n_action_floats)]"""
        raise NotLowerableError(
            f"decoder {type(self).__name__} has no device program (define device_ops())")

    def decode(self, ctx, action):
        raise DeviceOnlyError("Decoder.decode runs inside the fused step kernel")

    def chain(self, others: Iterable["Decoder"]) -> "ChainedDecoder":
        return ChainedDecoder(flatten([self, others]))

    def reset(self):
        pass

    def flat_dim(self) -> int:
        return sum(n for _, n in self.device_ops())

    def __repr__(self) -> str:
        return repr(self.action_space)

    def __str__(self) -> str:
        return str(self.action_space)


class EmptyDecoder(Decoder):
    """Takes an action and returns no messages (decoders.py:54-62)."""

    @property
    def action_space(self):
        return spaces.Box(-np.inf, np.inf, (1,))

    def device_ops(self):
        return [(OP_NO_MESSAGES, 1)]


class ChainedDecoder(Decoder):
    def __init__(self, decoders: Iterable[Decoder]):
        self.decoders: List[Decoder] = flatten(decoders)

    @property
    def action_space(self):
        return spaces.Tuple(tuple(d.action_space for d in self.decoders))

    def device_ops(self):
        return [op for d in self.decoders for op in d.device_ops()]

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
        return spaces.Dict({k: d.action_space for k, d in self.decoders.items()})

    def device_ops(self):
        return [op for d in self.decoders.values() for op in d.device_ops()]

    def reset(self):
        for d in self.decoders.values():
            d.reset()
