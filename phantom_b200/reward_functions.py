"""Reward functions (reference: phantom/reward_functions.py:6-38): a RewardFunction declares
the device program computing an agent's scalar reward.  Built in: `Constant`."""
from __future__ import annotations

from .errors import DeviceOnlyError, NotLowerableError

OP_CONST = 0


class RewardFunction:
    def device_op(self) -> tuple:
        """(opcode, value)"""
        raise NotLowerableError(
            f"reward function {type(self).__name__} has no device program (define device_op())")

    def reward(self, ctx) -> float:
        raise DeviceOnlyError("RewardFunction.reward runs inside the fused step kernel")

    def reset(self):
        pass


class Constant(RewardFunction):
    """Always returns `value` (reward_functions.py:26-38)."""

    def __init__(self, value: float = 0.0) -> None:
        self.value = value

    def device_op(self):
        return (OP_CONST, float(self.value))
