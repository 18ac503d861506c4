"""StackelbergEnv (reference: phantom/stackelberg.py:12-196): leaders act on odd steps,
followers on even steps; the turn logic and the reward cache run on the device
(PHX_ENV_STACKELBERG)."""
from __future__ import annotations

from typing import Sequence

from .env import PhantomEnv
from .network import Network
from .types import AgentID


class StackelbergEnv(PhantomEnv):
    def __init__(self, num_steps: int, network: Network, leader_agents: Sequence[AgentID],
                 follower_agents: Sequence[AgentID], env_supertype=None,
                 agent_supertypes=None, **batch_kwargs) -> None:
        super().__init__(num_steps, network, env_supertype, agent_supertypes, **batch_kwargs)
        for aid in list(leader_agents) + list(follower_agents):
            assert aid in network.agent_ids, f"Agent '{aid}' not in network"
        for aid in leader_agents:
            assert aid not in follower_agents, f"Agent '{aid}' not in network"
        self.leader_agents = leader_agents
        self.follower_agents = follower_agents
