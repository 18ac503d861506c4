"""Policy base class (reference: phantom/policy.py:7-36): a fixed / pre-trained policy that maps
one agent's observation to its action.  With a batched env (`num_envs > 1`) `compute_action`
receives the agent's observations of ALL envs at once, float32 `[E, obs_dim]`, and returns
`[E, act_dim]` (or `[E]`)."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any


class Policy(ABC):
    def __init__(self, observation_space, action_space) -> None:
        self.observation_space = observation_space
        self.action_space = action_space

    @abstractmethod
    def compute_action(self, observation: Any) -> Any:
        raise NotImplementedError
