"""SingleAgentEnvAdapter (reference: phantom/env_wrappers.py:23-196): a single-agent, gym-style
view of a multi-agent env -- one agent is driven by the caller, every other acting agent by a
fixed policy.  Same constructor, checks, properties and return tuples as the reference.  The
wrapped env steps on the GPU like any other; with `env_config={"num_envs": E, ...}` the adapter
is a VECTOR env: `step(action)` takes the selected agent's actions for all E envs
(`[E, act_dim]`) and returns `(obs [E, obs_dim], reward [E], terminated [E], truncated [E], {})`
as numpy arrays, and the other agents' policies see `[E, obs_dim]` observation batches.
"""
from __future__ import annotations

from typing import Any, Dict, List, Mapping, Optional, Tuple, Type

import numpy as np

from .agents import Agent
from .env import PhantomEnv
from .policy import Policy
from .types import AgentID


class SingleAgentEnvAdapter:
    def __init__(self, env_class: Type[PhantomEnv], agent_id: AgentID,
                 other_policies: Mapping[AgentID, Tuple[Type[Policy], Mapping[str, Any]]],
                 env_config: Optional[Mapping[str, Any]] = None) -> None:
        self._env = env_class(**(env_config or {}))
        if agent_id not in self._env.agent_ids:
            raise ValueError(
                f"Selected agent '{agent_id}' of SingleAgentEnvAdapter not found in underlying "
                f"env '{env_class.__name__}'")
        if agent_id in other_policies:
            raise ValueError(
                f"Selected agent '{agent_id}' of SingleAgentEnvAdapter found in agent ID to "
                "policy mapping")
        policies = list(other_policies.keys()) + [agent_id]
        for agent in self._env.agents.values():
            if getattr(agent, "action_space", None) is not None and agent.id not in policies:
                raise ValueError(
                    f"Agent '{agent.id}' has not been defined a policy via the 'other_policies' "
                    "parameter of SingleAgentEnvAdapter")
        self._agent_id = agent_id
        self._other_policies = {
            aid: policy_class(self._env[aid].observation_space, self._env[aid].action_space,
                              **policy_config)
            for aid, (policy_class, policy_config) in other_policies.items()
        }
        self._batched = self._env.num_envs > 1
        ids = self._env.strategic_agent_ids
        self._sidx = {aid: i for i, aid in enumerate(ids)}
        self._observations: Dict[AgentID, Any] = {}
        self._obs = self._obs_mask = None  # batched: last [E,S,O] / [E,S]
        self.reset()

    # ------------------------------------------------------------- reference properties
    @property
    def active_agent(self) -> AgentID:
        return self._agent_id

    @property
    def agents(self) -> Dict[AgentID, Agent]:
        return self._env.agents

    @property
    def agent_ids(self) -> List[AgentID]:
        return self._env.agent_ids

    @property
    def n_agents(self) -> int:
        return self._env.n_agents

    @property
    def current_step(self):
        return self._env.current_step

    @property
    def action_space(self):
        return self._env[self._agent_id].action_space

    @property
    def observation_space(self):
        return self._env[self._agent_id].observation_space

    @property
    def unwrapped(self) -> PhantomEnv:
        return self._env

    # ---------------------------------------------------------------------------- step
    def step(self, action):
        if not self._batched:
            actions = {aid: policy.compute_action(self._observations[aid])
                       for aid, policy in self._other_policies.items()
                       if aid in self._observations}
            actions[self._agent_id] = action
            step = self._env.step(actions)
            self._observations = step.observations
            me = self._agent_id
            return (step.observations.get(me), step.rewards.get(me), step.terminations.get(me),
                    step.truncations.get(me), step.infos.get(me, {}))
        env, E = self._env, self._env.num_envs
        S, A = len(self._sidx), env.spec.act_dim
        acts = np.zeros((E, S, A), np.float32)
        mask = np.zeros((E, S), np.uint8)
        obs, om = self._obs, self._obs_mask
        for aid, policy in self._other_policies.items():
            i = self._sidx[aid]
            O = env._agent_obs_dim(env[aid])
            a = np.asarray(policy.compute_action(obs[:, i, :O]), np.float32).reshape(E, -1)
            acts[:, i, : a.shape[1]] = a
            mask[:, i] = om[:, i]  # only agents that observed act (env.py:330-333 otherwise)
        i = self._sidx[self._agent_id]
        a = np.asarray(action, np.float32).reshape(E, -1)
        acts[:, i, : a.shape[1]] = a
        mask[:, i] = 1
        out = env.step_batch(acts, mask)
        self._obs = out.observations.cpu().numpy()
        self._obs_mask = out.obs_mask.cpu().numpy()
        O = env._agent_obs_dim(env[self._agent_id])
        return (self._obs[:, i, :O], out.rewards.cpu().numpy()[:, i],
                out.terminations.cpu().numpy()[:, i] == 1, out.truncations.cpu().numpy()[:, i] == 1, {})

    def reset(self):
        if not self._batched:
            self._observations, infos = self._env.reset()
            return self._observations.get(self._agent_id), infos
        obs, om = self._env.reset_batch()
        self._obs, self._obs_mask = obs.cpu().numpy(), om.cpu().numpy()
        i = self._sidx[self._agent_id]
        return self._obs[:, i, : self._env._agent_obs_dim(self._env[self._agent_id])], {}

    def close(self) -> None:
        self._env.close()
