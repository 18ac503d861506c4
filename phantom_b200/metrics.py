"""Metrics (reference: phantom/metrics.py:38-370).

Same classes and constructor signatures as the reference.  `extract(env)` reads agent / env
attributes exactly like the reference does (`_rgetattr`); on a device env those attributes are
state columns in HBM (`agents.device_column`), so for `num_envs == 1` the value is the
reference's scalar and for a batch it is an array [E].  `extract_over_envs(env, op)` reduces a
`SimpleAgentMetric` over the batch ON THE DEVICE (`phx_reduce_field`), which is what replaces
the per-replica Python loop of `RLlibMetricLogger.on_episode_step`
(phantom/utils/rllib/train.py:294-297).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from functools import reduce as _reduce
from typing import Callable, DefaultDict, Dict, Generic, Iterable, List, Optional, Sequence, TypeVar

import numpy as np

from .env import PhantomEnv
from .fsm import FiniteStateMachineEnv, FSMStage

MetricValue = TypeVar("MetricValue")


class NotRecorded:
    _instance = None

    def __new__(cls):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    def __repr__(self) -> str:
        return "<NotRecorded>"


not_recorded = NotRecorded()


class Metric(Generic[MetricValue], ABC):
    def __init__(self, fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        self.fsm_stages = fsm_stages
        self.description = description

    @abstractmethod
    def extract(self, env: PhantomEnv) -> MetricValue:
        raise NotImplementedError

    def reduce(self, values: Sequence[MetricValue], mode: str) -> MetricValue:
        return values[-1]


class LambdaMetric(Metric):
    def __init__(self, extract_fn: Callable[[PhantomEnv], MetricValue],
                 train_reduce_fn: Callable[[Sequence[MetricValue]], MetricValue],
                 eval_reduce_fn: Callable[[Sequence[MetricValue]], MetricValue],
                 fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        self.extract_fn = extract_fn
        self.train_reduce_fn = train_reduce_fn
        self.eval_reduce_fn = eval_reduce_fn
        super().__init__(fsm_stages, description)

    def extract(self, env: PhantomEnv) -> MetricValue:
        return self.extract_fn(env)

    def reduce(self, values, mode):
        if mode == "train":
            return self.train_reduce_fn(values)
        if mode == "evaluate":
            return self.eval_reduce_fn(values)
        raise ValueError(f"Unknown mode: {mode}")


class SimpleMetric(Metric, ABC):
    def __init__(self, train_reduce_action: str = "mean", eval_reduce_action: str = "none",
                 fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        if train_reduce_action not in ("last", "mean", "sum"):
            raise ValueError(
                f"train_reduce_action field of {self.__class__} metric must be one of: 'last', "
                f"'mean' or 'sum'. Got '{train_reduce_action}'.")
        if eval_reduce_action not in ("last", "mean", "sum", "none"):
            raise ValueError(
                f"eval_reduce_action field of {self.__class__} metric class must be one of: "
                f"'last', 'mean', 'sum' or 'none'. Got '{eval_reduce_action}'.")
        self.train_reduce_action = train_reduce_action
        self.eval_reduce_action = eval_reduce_action
        super().__init__(fsm_stages, description)

    def reduce(self, values, mode):
        action = self.train_reduce_action if mode == "train" else self.eval_reduce_action
        if action == "none":
            return np.array(values)
        if self.fsm_stages is not None:
            values = [v for v in values if v is not not_recorded]
        if action == "last":
            return values[-1] if len(values) > 0 else None
        if action == "mean":
            return np.mean(values, axis=0)
        if action == "sum":
            return np.sum(values, axis=0)
        raise ValueError


def _rgetattr(obj, attr, *args):
    return _reduce(lambda o, a: getattr(o, a, *args), [obj] + attr.split("."))


class SimpleAgentMetric(SimpleMetric):
    def __init__(self, agent_id: str, agent_property: str, train_reduce_action: str = "mean",
                 eval_reduce_action: str = "none",
                 fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        self.agent_id = agent_id
        self.agent_property = agent_property
        super().__init__(train_reduce_action, eval_reduce_action, fsm_stages, description)

    def extract(self, env: PhantomEnv):
        return _rgetattr(env.agents[self.agent_id], self.agent_property)

    def extract_over_envs(self, env: PhantomEnv, op: str = "mean") -> float:
        """The metric reduced over all envs of the batch on the device ('mean', 'sum', 'min',
        'max'); only for properties that are device state columns."""
        from .agents import _DeviceColumn

        agent = env.agents[self.agent_id]
        col = getattr(type(agent), self.agent_property, None)
        if not isinstance(col, _DeviceColumn):
            raise TypeError(f"'{self.agent_property}' is not a device state column")
        return env.reduce_agent_column(agent, col.word)[op]


class SimpleEnvMetric(SimpleMetric):
    def __init__(self, env_property: str, train_reduce_action: str = "mean",
                 eval_reduce_action: str = "none",
                 fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        self.env_property = env_property
        super().__init__(train_reduce_action, eval_reduce_action, fsm_stages, description)

    def extract(self, env: PhantomEnv):
        return _rgetattr(env, self.env_property)


class AggregatedAgentMetric(SimpleMetric):
    def __init__(self, agent_ids: Iterable[str], agent_property: str,
                 group_reduce_action: str = "mean", train_reduce_action: str = "mean",
                 eval_reduce_action: str = "none",
                 fsm_stages: Optional[Sequence[FSMStage]] = None,
                 description: Optional[str] = None) -> None:
        if group_reduce_action not in ["min", "max", "mean", "sum"]:
            raise ValueError("group_reduce_action field of SimpleMetric class must be one of: "
                             "'min', 'max', 'mean' or 'sum'.")
        self.agent_ids = agent_ids
        self.agent_property = agent_property
        self.group_reduce_action = group_reduce_action
        super().__init__(train_reduce_action, eval_reduce_action, fsm_stages, description)

    def extract(self, env: PhantomEnv):
        values = [_rgetattr(env.agents[a], self.agent_property) for a in self.agent_ids]
        fn = {"min": np.min, "max": np.max, "mean": np.mean, "sum": np.sum}[self.group_reduce_action]
        return fn(values, axis=0)  # over the agent group; per env for a batch


def logging_helper(env: PhantomEnv, metrics: Dict[str, Metric],
                   metric_values: DefaultDict[str, List[float]]) -> None:
    """metrics.py:355-370: record every metric whose FSM-stage filter matches the env's
    current stage.  A batch of FSM envs need not be in lock step (handler-driven StageRule
    transitions, auto-reset): when the stages differ across the batch the value is recorded per
    env, NaN where that env's stage is filtered out (and `not_recorded` if no env matches)."""
    for metric_id, metric in metrics.items():
        stage = None
        env_ok = None
        if isinstance(env, FiniteStateMachineEnv) and metric.fsm_stages is not None:
            stage = env.current_stage
            if not isinstance(stage, (str, int)) and np.ndim(stage) > 0:
                col = np.asarray(stage).ravel()
                ids = list(env._stages)
                allowed = [i for i, sid in enumerate(ids) if sid in metric.fsm_stages]
                env_ok = np.isin(col, allowed)
                stage = ids[int(col[0])] if (col == col[0]).all() else None
        if env_ok is not None and stage is None:  # stages differ across the batch
            if not env_ok.any():
                value = not_recorded
            else:
                value = np.array(metric.extract(env), dtype=np.float64, copy=True)
                value = np.where(env_ok.reshape((-1,) + (1,) * (value.ndim - 1)), value, np.nan)
        elif stage is None or stage in metric.fsm_stages:
            value = metric.extract(env)
        else:
            value = not_recorded
        metric_values[metric_id].append(value)
