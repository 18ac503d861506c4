"""Message payload declarations (reference: phantom/message.py:10-53).

On the device a message is the record {sender slot, receiver slot, payload type id,
payload words}; the Python classes below only *declare* payload types: their field list
and the sender / receiver agent-class whitelists that phantom/network.py:297-331 enforces.
`phantom_b200.lowering` turns the whitelists into per-slot bitmasks.
"""
from __future__ import annotations

import dataclasses
from typing import Any, Generic, List, Optional, TypeVar

from .types import AgentID

T = TypeVar("T")


@dataclasses.dataclass(frozen=True)
class MsgPayload:
    """Deprecated payload base class, kept for source compatibility."""


def _names(arg) -> Optional[List[str]]:
    if arg is None:
        return None
    seq = arg if isinstance(arg, list) else [arg]
    return [x.__name__ if isinstance(x, type) else x for x in seq]


def msg_payload(sender_type=None, receiver_type=None):
    """Declare a payload type; arguments are agent classes / class names / lists / None."""

    def wrap(cls):
        cls._sender_types = _names(sender_type)
        cls._receiver_types = _names(receiver_type)
        return dataclasses.dataclass(frozen=True)(cls)

    return wrap


@dataclasses.dataclass(frozen=True)
class Message(Generic[T]):
    """A routed message; produced by `PhantomEnv.tracked_messages()` from the device trace."""

    sender_id: AgentID
    receiver_id: AgentID
    payload: Any
