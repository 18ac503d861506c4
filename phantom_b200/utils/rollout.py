"""Rollout records (reference: phantom/utils/rollout.py:23-341).

`AgentStep`, `Step` and `Rollout` are the reference's frozen records with the same fields and
helper methods, `rollouts_to_dataframe` / `rollouts_to_jsonl` / `RolloutJSONEncoder` the same
exporters.  What differs is where the records come from: the reference's rollout utility steps a
Python list of env objects and appends one `Step` per env and step
(phantom/utils/rllib/rollout.py:289-363).  Here ONE `phx_rollout` launch advances every env by
T steps on the device; `rollouts_from_batch` turns its output planes (and, with
`BatchResolver(enable_tracking=True)`, the device message trace of every step) into the
per-env records afterwards, only for the envs somebody asks for.
"""
from __future__ import annotations

import io
import json
from collections import Counter
from dataclasses import asdict, dataclass
from typing import Any, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from ..message import Message
from ..types import AgentID, StageID


@dataclass(frozen=True)
class AgentStep:
    """Describes a step taken by a single agent in an episode."""

    i: int
    observation: Optional[Any]
    reward: Optional[float]
    done: bool
    info: Optional[Dict[str, Any]]
    action: Optional[Any]
    stage: Optional[StageID] = None


@dataclass(frozen=True)
class Step:
    """Describes a step taken in an episode."""

    i: int
    observations: Dict[AgentID, Any]
    rewards: Dict[AgentID, float]
    terminations: Dict[AgentID, bool]
    truncations: Dict[AgentID, bool]
    infos: Dict[AgentID, Dict[str, Any]]
    actions: Dict[AgentID, Any]
    messages: Optional[List[Message]] = None
    stage: Optional[StageID] = None


@dataclass(frozen=True)
class Rollout:
    rollout_id: int
    repeat_id: int
    env_config: Mapping[str, Any]
    rollout_params: Dict[str, Any]
    steps: List[Step]
    metrics: Dict[str, np.ndarray]

    def _select(self, field: str, agent_id: AgentID, drop_nones: bool,
                stages: Optional[Iterable[StageID]], none_values: bool = False):
        out = []
        for step in self.steps:
            table = getattr(step, field)
            present = agent_id in table and not (none_values and table[agent_id] is None)
            if (drop_nones is False or present) and (stages is None or step.stage in stages):
                out.append(table.get(agent_id, None))
        return out

    def observations_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                               stages: Optional[Iterable[StageID]] = None) -> List[Optional[Any]]:
        return self._select("observations", agent_id, drop_nones, stages)

    def rewards_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                          stages: Optional[Iterable[StageID]] = None) -> List[Optional[float]]:
        return self._select("rewards", agent_id, drop_nones, stages, none_values=True)

    def terminations_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                               stages: Optional[Iterable[StageID]] = None) -> List[Optional[bool]]:
        return self._select("terminations", agent_id, drop_nones, stages)

    def truncations_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                              stages: Optional[Iterable[StageID]] = None) -> List[Optional[bool]]:
        return self._select("truncations", agent_id, drop_nones, stages)

    def infos_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                        stages: Optional[Iterable[StageID]] = None) -> List[Optional[Dict[str, Any]]]:
        return self._select("infos", agent_id, drop_nones, stages)

    def actions_for_agent(self, agent_id: AgentID, drop_nones: bool = False,
                          stages: Optional[Iterable[StageID]] = None) -> List[Optional[Any]]:
        return self._select("actions", agent_id, drop_nones, stages)

    def steps_for_agent(self, agent_id: AgentID,
                        stages: Optional[Iterable[StageID]] = None) -> List[AgentStep]:
        steps = self.steps if stages is None else [s for s in self.steps if s.stage in stages]
        # (the reference passes terminations AND truncations positionally into AgentStep's single
        # `done` field, rollout.py:197-208, which raises TypeError; `done` here is their OR)
        def done(s):
            te, tr = s.terminations.get(agent_id, None), s.truncations.get(agent_id, None)
            return None if te is None and tr is None else bool(te) or bool(tr)

        return [AgentStep(s.i, s.observations.get(agent_id, None), s.rewards.get(agent_id, None),
                          done(s), s.infos.get(agent_id, None), s.actions.get(agent_id, None), s.stage)
                for s in steps]

    def count_actions(self, stages: Optional[Iterable[StageID]] = None) -> List[Tuple[Any, int]]:
        acts = (a for s in self.steps if stages is None or s.stage in stages
                for a in s.actions.values())
        return Counter(acts).most_common()

    def count_agent_actions(self, agent_id: AgentID,
                            stages: Optional[Iterable[StageID]] = None) -> List[Tuple[Any, int]]:
        acts = (s.actions.get(agent_id, None) for s in self.steps
                if stages is None or s.stage in stages)
        return Counter(acts).most_common()

    def __getitem__(self, index: int):
        """Returns a step for a given index in the episode."""
        try:
            return self.steps[index]
        except KeyError:
            raise KeyError(f"Index {index} not valid for trajectory")


def rollouts_to_dataframe(rollouts: Iterable[Rollout], avg_over_repeats: bool = True,
                          index_value_precision: Optional[int] = None):
    """MultiIndex DataFrame with the rollout params as index and the metrics as columns
    (rollout.py:260-299)."""
    import pandas as pd

    rows = [(r.rollout_params, r.metrics) for r in rollouts]
    index_cols = list(rows[0][0].keys())
    df = pd.DataFrame([{**params, **metrics} for params, metrics in rows])
    if index_value_precision is not None:
        for col in index_cols:
            df[col] = df[col].round(index_value_precision).astype(str)
    if len(index_cols) > 0:
        if avg_over_repeats:
            df = df.groupby(index_cols).mean().reset_index()
        df = df.set_index(index_cols)
    return df


class RolloutJSONEncoder(json.JSONEncoder):
    def default(self, o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, np.bool_):
            return bool(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.number):
            return int(o)
        if isinstance(o, (Rollout, Step)):
            return asdict(o)
        return json.JSONEncoder.default(self, o)


def rollouts_to_jsonl(rollouts: Iterable[Rollout], file_obj: io.TextIOBase,
                      human_readable: bool = False) -> None:
    """One JSON document per rollout and line (rollout.py:302-322)."""
    for rollout in rollouts:
        json.dump(rollout, file_obj, indent=2 if human_readable else None, cls=RolloutJSONEncoder)
        file_obj.write("\n")
        file_obj.flush()


# ----------------------------------------------------------------- device batch -> records
def _host(x):
    return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)


def rollouts_from_batch(env, actions, out, action_mask=None, env_indices: Optional[Sequence[int]] = None,
                        record_messages: bool = False, env_config: Optional[Mapping[str, Any]] = None,
                        rollout_params: Optional[Dict[str, Any]] = None,
                        metrics: Optional[Dict[str, Any]] = None, rollout_id_base: int = 0,
                        stages=None) -> List[Rollout]:
    """Per-env `Rollout` records of ONE T-step device rollout.

    env            the batched env that was stepped
    actions        [T,E,S,A] what rollout_batch received (tensor or array)
    out            the BatchStep rollout_batch returned (leading T axis)
    action_mask    [T,E,S] or None: 0 = the agent was absent from the `actions` mapping
    env_indices    which envs to materialise (default: all)
    record_messages  attach Resolver.tracked_messages of every step (needs a handle created with
                   BatchResolver(enable_tracking=True): the device records them during the
                   rollout, phx_get_trace_step)
    stages         FSM envs: [T, E] stage index BEFORE each step, e.g. collected with
                   env.field(FIELD_STAGE) by a stepping caller; None leaves Step.stage = None
    `Step.i` counts from 0 as in the reference's rollout loop; the observation of step i is the
    one returned BY step i (rollout.py's `new_observations`).
    """
    A = _host(actions)
    M = None if action_mask is None else _host(action_mask)
    planes = [_host(p) for p in out]
    T, E = A.shape[0], env.num_envs
    idx = list(range(E)) if env_indices is None else [int(i) for i in env_indices]
    ids = env.strategic_agent_ids
    agents = env.strategic_agents
    stage_ids = list(getattr(env, "_stages", {})) or None
    rollouts = []
    for e in idx:
        msgs = env.rollout_messages(e) if record_messages else None
        steps = []
        for t in range(T):
            view = env._step_from_host(*[p[t, e] for p in planes])
            acts = {}
            for s, agent in enumerate(agents):
                if M is None or M[t, e, s]:
                    a = A[t, e, s]
                    acts[agent.id] = (int(round(float(a[0])))
                                      if hasattr(getattr(agent, "action_space", None), "n") else a.copy())
            stage = None
            if stages is not None and stage_ids is not None:
                stage = stage_ids[int(_host(stages)[t, e])]
            steps.append(Step(t, view.observations, view.rewards, view.terminations,
                              view.truncations, view.infos, acts,
                              None if msgs is None else msgs[t], stage))
        metric_e = {k: (np.asarray(v)[e] if np.ndim(v) > 0 else v) for k, v in (metrics or {}).items()}
        rollouts.append(Rollout(rollout_id_base + e, 0, dict(env_config or {}),
                                dict(rollout_params or {}), steps, metric_e))
    return rollouts
