"""Vector-env façade over a batched device env.

The reference's callers step a Python list of env objects and batch policy inference per agent
id (`phantom/utils/rllib/rollout.py:289-363`: `vec_envs = [env_class(**cfg) ...]`, then for each
agent id one batched `compute_actions`, then `[env.step(a) for env, a in zip(...)]`), or wrap
one env per RLlib worker (`utils/rllib/wrapper.py:10-57`).  `VectorEnv` gives those callers the
same information with tensors:

    vec = VectorEnv(SupplyChainEnv(num_envs=4096))
    obs = vec.reset()                         # {agent_id: tensor [E, obs_dim_of_agent]}
    step = vec.step({"SHOP": actions})        # actions: {agent_id: tensor [E, act_dim]}
    step.observations["SHOP"], step.rewards["SHOP"], step.obs_mask["SHOP"] ...
    step.env(17)                              # lazy PhantomEnv.Step dict view of sub-env 17

Agents absent from the `actions` dict fall back to `generate_messages()` in every env, exactly
like a missing key in the reference's `actions` mapping (env.py:330-333); per-env absence is
expressed with `action_mask={agent_id: bool tensor [E]}`.
"""
from __future__ import annotations

from typing import Any, Dict, Mapping, NamedTuple, Optional

import numpy as np

from .env import BatchStep, PhantomEnv
from .types import AgentID


class VectorStep(NamedTuple):
    observations: Dict[AgentID, Any]   # float32 [E, O_a]   (valid where obs_mask)
    obs_mask: Dict[AgentID, Any]       # bool    [E]       agent is a key of Step.observations
    rewards: Dict[AgentID, Any]        # float32 [E]
    reward_mask: Dict[AgentID, Any]    # uint8   [E]       0 absent, 1 value, 2 None
    terminations: Dict[AgentID, Any]   # uint8   [E]       255 = key absent (already done)
    truncations: Dict[AgentID, Any]    # uint8   [E]
    all_terminated: Any                # bool [E]   terminations["__all__"]
    all_truncated: Any                 # bool [E]   truncations["__all__"]
    batch: BatchStep                   # the underlying [E,S,...] tensors
    owner: Any

    def env(self, index: int) -> PhantomEnv.Step:
        """PhantomEnv.Step of sub-env `index`, built on demand (copies one row to the host)."""
        host = [t[index].cpu().numpy() for t in self.batch]
        return self.owner.env._step_from_host(*host)


class VectorEnv:
    def __init__(self, env: PhantomEnv) -> None:
        self.env = env
        self.agent_ids = env.strategic_agent_ids
        self.num_envs = env.num_envs
        self._obs_dims = {a.id: env._agent_obs_dim(a) for a in env.strategic_agents}
        for a in env.strategic_agents:
            enc = getattr(a, "observation_encoder", None)
            if enc is not None:
                self._obs_dims[a.id] = enc.flat_dim()

    @property
    def observation_spaces(self) -> Dict[AgentID, Any]:
        return {a.id: a.observation_space for a in self.env.strategic_agents}

    @property
    def action_spaces(self) -> Dict[AgentID, Any]:
        return {a.id: a.action_space for a in self.env.strategic_agents}

    def _split(self, batch: BatchStep) -> VectorStep:
        obs, om, rew, rm, te, tr = {}, {}, {}, {}, {}, {}
        for s, aid in enumerate(self.agent_ids):
            obs[aid] = batch.observations[:, s, : self._obs_dims[aid]]
            om[aid] = batch.obs_mask[:, s].bool()
            rew[aid] = batch.rewards[:, s]
            rm[aid] = batch.reward_mask[:, s]
            te[aid] = batch.terminations[:, s]
            tr[aid] = batch.truncations[:, s]
        return VectorStep(obs, om, rew, rm, te, tr, batch.all_done[:, 0].bool(),
                          batch.all_done[:, 1].bool(), batch, self)

    def reset(self, env_mask=None) -> Dict[AgentID, Any]:
        """{agent_id: obs [E, O_a]} for the agents the reference's reset() returns."""
        obs, mask = self.env.reset_batch(env_mask)
        # an agent is returned if ANY reset env observed it: reset_batch zeroes the mask plane
        # and the kernel writes only the rows of the masked envs, and an agent's encode may
        # return None in some envs only (digital_ads); per-env presence is `reset_obs_mask`
        self.reset_obs_mask = {aid: mask[:, s].bool() for s, aid in enumerate(self.agent_ids)}
        mask_host = mask.any(0).cpu().numpy()
        return {aid: obs[:, s, : self._obs_dims[aid]]
                for s, aid in enumerate(self.agent_ids) if mask_host[s]}

    def step(self, actions: Mapping[AgentID, Any],
             action_mask: Optional[Mapping[AgentID, Any]] = None) -> VectorStep:
        torch = self.env._torch()
        dev = torch.device("cuda", self.env.device)
        E, S, A = self.num_envs, max(len(self.agent_ids), 1), self.env.spec.act_dim
        a = torch.zeros((E, S, A), dtype=torch.float32, device=dev)
        m = torch.zeros((E, S), dtype=torch.uint8, device=dev)
        for s, aid in enumerate(self.agent_ids):
            if aid in actions:
                v = torch.as_tensor(actions[aid], dtype=torch.float32, device=dev).reshape(E, -1)
                a[:, s, : v.shape[1]] = v[:, :A]
                m[:, s] = 1
                if action_mask is not None and aid in action_mask:
                    m[:, s] = torch.as_tensor(action_mask[aid], device=dev).to(torch.uint8)
        return self._split(self.env.step_batch(a, m))

    def close(self) -> None:
        self.env.close()
