"""FiniteStateMachineEnv (reference: phantom/fsm.py:12-380).

Host side: stage registration and the reference's validation errors.  The stage-gated step
loop, the reward / observation caches and the transition run on the device
(PHX_ENV_FSM).  Handler-less (deterministic) stages lower as they are; a stage WITH an env
handler (fsm.py:294-307) lowers when the handler is a `StageRule` -- the declarative form of
`[self.resolve_network()]; return A if <state> <cmp> <const> else B`, evaluated on the device
after the stage's message resolution.  An arbitrary Python handler raises NotLowerableError.
"""
from __future__ import annotations

from typing import Callable, Dict, Mapping, Optional, Sequence

import dataclasses

from .env import PhantomEnv
from .network import Network
from .types import AgentID, StageID
from .views import EnvView


@dataclasses.dataclass(frozen=True)
class FSMEnvView(EnvView):
    """reference: phantom/fsm.py:66-73"""

    stage: StageID


class FSMValidationError(Exception):
    pass


class FSMRuntimeError(Exception):
    pass


class StageRule:
    """Device-lowerable env stage handler (reference: the `handler` of `FSMStage`,
    phantom/fsm.py:33-63, called at fsm.py:294-302).

    Stands for the Python handler ::

        def handler(env):
            if resolve_network: env.resolve_network()       # fsm.py:280-283: a stage with a
            if <lhs> <cmp> <rhs> [and <also>]:              # handler is resolved only by it
                return then
            elif <terms of elifs[0]>: return <stage of elifs[0]>
            ...
            return otherwise

    lhs:  "always" | "step" (env.current_step) | ("agent", agent_id, column) where column is a
          device_column attribute name of that agent's class or a state word index |
          ("env", word) for families with env-level words.
    cmp:  one of "<", "<=", "==", "!=", ">=", ">".
    rhs:  an int32 constant, or another operand ("step" / ("agent", ...) / ("env", word)); for a
          float32 device column: a float constant that is exactly representable in float32, or
          another float32 column (the comparison is then made in float32, as between numpy
          float32 scalars).
    also: a second comparison `(lhs, cmp, rhs)` that must hold as well (AND).
    elifs: further branches `(terms, stage)` tried in order when the first does not hold, `terms`
          one comparison `(lhs, cmp, rhs)` or a list of up to TERMS of them (AND).  OR is two
          branches returning the same stage; up to BRANCHES branches in all.
    A returned stage outside the stage's `next_stages` faults the env with FSMRuntimeError
    (fsm.py:304-307).  The object is callable only so that it can sit where the reference
    expects a handler; calling it on the host raises DeviceOnlyError.
    """

    CMPS = ("<", "<=", "==", "!=", ">=", ">")
    BRANCHES, TERMS = 4, 2  # include/phx.h PHX_RULE_BRANCHES / PHX_RULE_TERMS

    def __init__(self, then: StageID, lhs="always", cmp: str = "==", rhs=0,
                 otherwise: Optional[StageID] = None, resolve_network: bool = True,
                 also=None, elifs=None) -> None:
        if lhs != "always" and otherwise is None:
            raise ValueError("StageRule: a conditional rule needs `otherwise`")
        if lhs == "always" and (also is not None or elifs):
            raise ValueError("StageRule: an unconditional rule has no further comparisons")
        first = [(lhs, cmp, rhs)] + ([] if also is None else [tuple(also)])
        self.branches = [(first, then)]
        for terms, stage in elifs or ():
            terms = [tuple(terms)] if isinstance(terms, tuple) and len(terms) == 3 and \
                isinstance(terms[1], str) and terms[1] in self.CMPS else [tuple(t) for t in terms]
            self.branches.append((terms, stage))
        if len(self.branches) > self.BRANCHES:
            raise ValueError(f"StageRule: at most {self.BRANCHES} branches")
        for terms, _ in self.branches:
            if not 1 <= len(terms) <= self.TERMS:
                raise ValueError(f"StageRule: 1..{self.TERMS} comparisons per branch")
            for t in terms:
                if len(t) != 3 or t[1] not in self.CMPS:
                    raise ValueError(f"StageRule: unknown comparison {t!r}")
                if t[0] == "always" and t is not first[0]:
                    raise ValueError("StageRule: 'always' only as the single unconditional rule")
        self.then, self.lhs, self.cmp, self.rhs = then, lhs, cmp, rhs
        self.otherwise = then if otherwise is None else otherwise
        self.resolve_network = bool(resolve_network)

    def __call__(self, *args, **kwargs):
        from .errors import DeviceOnlyError

        raise DeviceOnlyError("a StageRule is evaluated by the fused step kernel")


class FSMStage:
    def __init__(self, stage_id: StageID, acting_agents: Sequence[AgentID],
                 rewarded_agents: Optional[Sequence[AgentID]] = None,
                 next_stages: Optional[Sequence[StageID]] = None,
                 handler: Optional[Callable[[], StageID]] = None) -> None:
        self.id = stage_id
        self.acting_agents = acting_agents
        self.rewarded_agents = rewarded_agents
        self.next_stages = next_stages or []
        self.handler = handler

    def __call__(self, handler_fn):
        handler_fn._decorator = self
        self.handler = handler_fn
        return handler_fn


class FiniteStateMachineEnv(PhantomEnv):
    def __init__(self, num_steps: int, network: Network, initial_stage: StageID,
                 env_supertype=None, agent_supertypes=None,
                 stages: Optional[Sequence[FSMStage]] = None, **batch_kwargs) -> None:
        super().__init__(num_steps, network, env_supertype, agent_supertypes, **batch_kwargs)
        self._initial_stage = initial_stage
        self._stages: Dict[StageID, FSMStage] = {}
        self.previous_stage: Optional[StageID] = None
        for st in stages or []:
            self._stages.setdefault(st.id, st)
        for name in dir(type(self)):
            attr = getattr(type(self), name, None)
            if callable(attr) and hasattr(attr, "_decorator"):
                if attr._decorator.id in self._stages:
                    raise FSMValidationError(f"Found multiple stages with ID '{attr._decorator.id}'")
                self._stages[attr._decorator.id] = attr._decorator
        if not self._stages:
            raise FSMValidationError("No registered stages.")
        if initial_stage not in self._stages:
            raise FSMValidationError(f"Initial stage '{initial_stage}' is not a valid stage")
        for st in self._stages.values():
            for nxt in st.next_stages:
                if nxt not in self._stages:
                    raise FSMValidationError(
                        f"Next stage '{nxt}' given in stage '{st.id}' is not a valid stage")
        for st in self._stages.values():
            if len(st.next_stages) != 1 and st.handler is None:
                raise FSMValidationError(
                    f"Stage '{st.id}' without handler must have exactly one next stage "
                    f"(got {len(st.next_stages)})")

    @property
    def initial_stage(self) -> StageID:
        return self._initial_stage

    @property
    def current_stage(self):
        from . import _lib as L
        import numpy as np

        ids = list(self._stages)
        if not self.is_live:
            return self._initial_stage
        col = self.field(L.FIELD_STAGE, np.int32)
        return ids[int(col[0])] if self.num_envs == 1 else col

    def view(self, agent_views=None) -> FSMEnvView:
        step = self.current_step if self.num_envs == 1 else 0
        stage = self.current_stage if self.num_envs == 1 else self._initial_stage
        return FSMEnvView(step, step / self.num_steps, stage)

    def is_fsm_deterministic(self) -> bool:
        return all(len(s.next_stages) == 1 for s in self._stages.values())
