"""Resolvers (reference: phantom/resolvers.py:17-163).

The round loop itself is device code (csrc/phx_engine.cuh, phx_engine1.cuh); these classes carry the
resolver *configuration* into the spec: round_limit, tracking, shuffle.
"""
from __future__ import annotations

from typing import List, Optional

from .errors import DeviceOnlyError
from .message import Message


class Resolver:
    def __init__(self, enable_tracking: bool = False) -> None:
        self.enable_tracking = enable_tracking
        self._tracked_messages: List[Message] = []

    def push(self, message: Message) -> None:
        raise DeviceOnlyError("messages are pushed by the fused step kernel")

    def clear_tracked_messages(self) -> None:
        self._tracked_messages.clear()

    @property
    def tracked_messages(self) -> List[Message]:
        """Messages of the steps taken since the last clear (filled from the device trace
        after every step when tracking is enabled; num_envs == 1 only)."""
        return self._tracked_messages

    def reset(self) -> None:
        pass


class BatchResolver(Resolver):
    def __init__(self, enable_tracking: bool = False, round_limit: Optional[int] = None,
                 shuffle_batches: bool = False) -> None:
        super().__init__(enable_tracking)
        self.round_limit = round_limit
        self.shuffle_batches = shuffle_batches
