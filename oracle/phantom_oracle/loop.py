"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The three env step loops of the reference, restated.

Reference lines followed (all under /root/reference/phantom/):
  env.py:48-53      PhantomEnv.Step
  env.py:55-124     PhantomEnv.__init__ (sampler collection, initial agent.reset())
  env.py:126-183    properties, view, pre/post_message_resolution, resolve_network
  env.py:185-237    reset
  env.py:239-303    step
  env.py:308-318    is_terminated / is_truncated
  env.py:320-348    _handle_acting_agents / _make_ctxs
  fsm.py:12-63      FSMValidationError / FSMRuntimeError / FSMStage
  fsm.py:66-73      FSMEnvView
  fsm.py:104-193    FiniteStateMachineEnv.__init__ (registration + validation), view
  fsm.py:195-251    reset
  fsm.py:253-380    step
  stackelberg.py:30-50   StackelbergEnv.__init__
  stackelberg.py:53-109  reset
  stackelberg.py:111-196 step
"""
from __future__ import annotations

import dataclasses
from typing import Any, Callable, Dict, List, Mapping, NamedTuple, Optional, Sequence, Set, Tuple

from .core import (
    Agent,
    AgentID,
    AgentView,
    Context,
    EnvView,
    Sampler,
    StageID,
    StrategicAgent,
    Supertype,
)
from .net import Network


class PhantomEnv:
    class Step(NamedTuple):
        observations: Dict[AgentID, Any]
        rewards: Dict[AgentID, float]
        terminations: Dict[AgentID, bool]
        truncations: Dict[AgentID, bool]
        infos: Dict[AgentID, Any]

    def __init__(
        self,
        num_steps: int,
        network: Optional[Network] = None,
        env_supertype: Optional[Supertype] = None,
        agent_supertypes: Optional[Mapping[AgentID, Supertype]] = None,
    ):
        self.network = network or Network()
        self._current_step = 0
        self.num_steps = num_steps
        self.env_supertype: Optional[Supertype] = None
        self.env_type: Optional[Supertype] = None
        self._terminations: Set[AgentID] = set()
        self._truncations: Set[AgentID] = set()
        self._ctxs: Dict[AgentID, Context] = {}
        self._samplers: List[Sampler] = []

        def collect(st):
            st._managed = True
            for value in st.__dict__.values():
                # identity-free `in`: uses ==, as env.py:95-96 does
                if isinstance(value, Sampler) and value not in self._samplers:
                    self._samplers.append(value)

        if env_supertype is not None:
            if isinstance(env_supertype, dict):
                env_supertype = self.Supertype(**env_supertype)
            else:
                assert isinstance(env_supertype, self.Supertype)
            collect(env_supertype)
            self.env_supertype = env_supertype

        if agent_supertypes is not None:
            for aid, st in agent_supertypes.items():
                if isinstance(st, dict):
                    st = self.agents[aid].Supertype(**st)
                collect(st)
                self.network.agents[aid].supertype = st

        for s in self._samplers:
            s.sample()
        for agent in self.agents.values():
            agent.reset()

    # ------------------------------------------------------------------- accessors
    @property
    def current_step(self) -> int:
        return self._current_step

    @property
    def n_agents(self) -> int:
        return len(self.agent_ids)

    @property
    def agents(self) -> Dict[AgentID, Agent]:
        return self.network.agents

    @property
    def agent_ids(self) -> List[AgentID]:
        return list(self.network.agent_ids)

    @property
    def strategic_agents(self) -> List[StrategicAgent]:
        return [a for a in self.agents.values() if isinstance(a, StrategicAgent)]

    @property
    def non_strategic_agents(self) -> List[Agent]:
        return [a for a in self.agents.values() if not isinstance(a, StrategicAgent)]

    @property
    def strategic_agent_ids(self) -> List[AgentID]:
        return [a.id for a in self.strategic_agents]

    @property
    def non_strategic_agent_ids(self) -> List[AgentID]:
        return [a.id for a in self.non_strategic_agents]

    def __getitem__(self, agent_id: AgentID) -> Agent:
        return self.network[agent_id]

    # ----------------------------------------------------------------------- hooks
    def view(self, agent_views: Dict[AgentID, AgentView]) -> EnvView:
        return EnvView(self.current_step, self.current_step / self.num_steps)

    def pre_message_resolution(self) -> None:
        for ctx in self._ctxs.values():
            ctx.agent.pre_message_resolution(ctx)

    def post_message_resolution(self) -> None:
        for ctx in self._ctxs.values():
            ctx.agent.post_message_resolution(ctx)

    def resolve_network(self) -> None:
        self.pre_message_resolution()
        self.network.resolve(self._ctxs)
        self.post_message_resolution()

    # ------------------------------------------------------------------ reset/step
    def _reset_common(self) -> None:
        self._current_step = 0
        for s in self._samplers:
            s.sample()
        if self.env_supertype is not None:
            self.env_type = self.env_supertype.sample()
        self.network.reset()
        self._terminations = set()
        self._truncations = set()

    def _initial_obs(self, agent_ids: Sequence[AgentID]):
        self._make_ctxs(agent_ids)
        obs = {c.agent.id: c.agent.encode_observation(c) for c in self._ctxs.values()}
        return {k: v for k, v in obs.items() if v is not None}, {}

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        self._reset_common()
        return self._initial_obs(self.strategic_agent_ids)

    def _done(self, aid: AgentID) -> bool:
        return aid in self._terminations or aid in self._truncations

    def _flag_agent(self, aid, ctx, terminations, truncations) -> None:
        terminations[aid] = ctx.agent.is_terminated(ctx)
        truncations[aid] = ctx.agent.is_truncated(ctx)
        if terminations[aid]:
            self._terminations.add(aid)
        if truncations[aid]:
            self._truncations.add(aid)

    def step(self, actions: Mapping[AgentID, Any]) -> "PhantomEnv.Step":
        self._current_step += 1
        self._make_ctxs(self.agent_ids)
        self._handle_acting_agents(self.agent_ids, actions)
        self.resolve_network()

        observations, rewards, terminations, truncations, infos = {}, {}, {}, {}, {}
        for aid in self.strategic_agent_ids:
            if self._done(aid):
                continue
            ctx = self._ctxs[aid]
            obs = ctx.agent.encode_observation(ctx)
            if obs is not None:  # env.py:281-284: obs, info, reward travel together
                observations[aid] = obs
                infos[aid] = ctx.agent.collect_infos(ctx)
                rewards[aid] = ctx.agent.compute_reward(ctx)
            self._flag_agent(aid, ctx, terminations, truncations)

        terminations["__all__"] = self.is_terminated()
        truncations["__all__"] = self.is_truncated()
        return self.Step(observations, rewards, terminations, truncations, infos)

    def render(self) -> None:
        return None

    def is_terminated(self) -> bool:
        return len(self._terminations) == len(self.strategic_agents)

    def is_truncated(self) -> bool:
        at_max = self.num_steps is not None and self.current_step == self.num_steps
        return at_max or len(self._truncations) == len(self.strategic_agents)

    def _handle_acting_agents(self, agent_ids: Sequence[AgentID], actions) -> None:
        for aid in agent_ids:
            if self._done(aid):
                continue
            ctx = self._ctxs[aid]
            if aid in actions:
                messages = ctx.agent.decode_action(ctx, actions[aid]) or []
            else:
                messages = ctx.agent.generate_messages(ctx) or []
            for receiver_id, payload in messages:
                self.network.send(aid, receiver_id, payload)

    def _make_ctxs(self, agent_ids: Sequence[AgentID]) -> None:
        env_view = self.view({aid: a.view() for aid, a in self.agents.items()})
        self._ctxs = {
            aid: self.network.context_for(aid, env_view)
            for aid in agent_ids
            if not self._done(aid)
        }


# ================================================================================= FSM
class FSMValidationError(Exception):
    pass


class FSMRuntimeError(Exception):
    pass


class FSMStage:
    def __init__(
        self,
        stage_id: StageID,
        acting_agents: Sequence[AgentID],
        rewarded_agents: Optional[Sequence[AgentID]] = None,
        next_stages: Optional[Sequence[StageID]] = None,
        handler: Optional[Callable[[], StageID]] = None,
    ):
        self.id = stage_id
        self.acting_agents = acting_agents
        self.rewarded_agents = rewarded_agents
        self.next_stages = next_stages or []
        self.handler = handler

    def __call__(self, handler_fn):
        handler_fn._decorator = self
        self.handler = handler_fn
        return handler_fn


@dataclasses.dataclass(frozen=True)
class FSMEnvView(EnvView):
    stage: StageID


class FiniteStateMachineEnv(PhantomEnv):
    def __init__(
        self,
        num_steps: int,
        network: Network,
        initial_stage: StageID,
        env_supertype: Optional[Supertype] = None,
        agent_supertypes: Optional[Mapping[AgentID, Supertype]] = None,
        stages: Optional[Sequence[FSMStage]] = None,
    ):
        super().__init__(num_steps, network, env_supertype, agent_supertypes)
        self._initial_stage = initial_stage
        self._rewards: Dict[AgentID, Optional[float]] = {}
        self._observations: Dict[AgentID, Any] = {}
        self._infos: Dict[AgentID, Dict[str, Any]] = {}
        self._stages: Dict[StageID, FSMStage] = {}
        self._current_stage = initial_stage
        self.previous_stage: Optional[StageID] = None

        for stage in stages or []:  # fsm.py:131-134: first registration wins
            self._stages.setdefault(stage.id, stage)
        for attr_name in dir(self):  # fsm.py:136-147: decorator-registered stages
            attr = getattr(self, attr_name)
            if callable(attr) and hasattr(attr, "_decorator"):
                if attr._decorator.id in self._stages:
                    raise FSMValidationError(
                        f"Found multiple stages with ID '{attr._decorator.id}'"
                    )
                self._stages[attr._decorator.id] = attr._decorator

        if not self._stages:
            raise FSMValidationError("No registered stages.")
        if self.initial_stage not in self._stages:
            raise FSMValidationError(f"Initial stage '{self.initial_stage}' is not a valid stage")
        for stage in self._stages.values():
            for nxt in stage.next_stages:
                if nxt not in self._stages:
                    raise FSMValidationError(
                        f"Next stage '{nxt}' given in stage '{stage.id}' is not a valid stage"
                    )
        for stage in self._stages.values():
            if len(stage.next_stages) != 1 and stage.handler is None:
                raise FSMValidationError(
                    f"Stage '{stage.id}' without handler must have exactly one next stage "
                    f"(got {len(stage.next_stages)})"
                )

    @property
    def initial_stage(self) -> StageID:
        return self._initial_stage

    @property
    def current_stage(self) -> StageID:
        return self._current_stage

    def is_fsm_deterministic(self) -> bool:
        return all(len(s.next_stages) == 1 for s in self._stages.values())

    def view(self, agent_views: Dict[AgentID, AgentView]) -> FSMEnvView:
        return FSMEnvView(
            self.current_step, self.current_step / self.num_steps, self.current_stage
        )

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        self._current_stage = self.initial_stage
        self._reset_common()
        self._rewards = {aid: None for aid in self.strategic_agent_ids}
        acting = self._stages[self.current_stage].acting_agents
        strategic = self.strategic_agent_ids
        return self._initial_obs([aid for aid in acting if aid in strategic])

    def step(self, actions: Mapping[AgentID, Any]) -> PhantomEnv.Step:
        self._current_step += 1
        self._make_ctxs(self.agent_ids)
        stage = self._stages[self.current_stage]
        self._handle_acting_agents(stage.acting_agents, actions)

        handler = stage.handler
        if handler is None:  # fsm.py:281-292
            self.resolve_network()
            if len(stage.next_stages) == 0:
                raise ValueError(
                    f"Current stage '{self.current_stage}' does not have an env handler "
                    "or a next stage defined"
                )
            next_stage = stage.next_stages[0]
        elif hasattr(handler, "__self__"):  # bound method given through `stages=`
            next_stage = handler()
        else:  # decorator-registered plain function
            next_stage = handler(self)

        if next_stage not in stage.next_stages:
            raise FSMRuntimeError(
                f"FiniteStateMachineEnv attempted invalid transition from "
                f"'{self.current_stage}' to {next_stage}"
            )

        observations, rewards, terminations, truncations, infos = {}, {}, {}, {}, {}
        if stage.rewarded_agents is None:  # fsm.py:315-320
            rewarded = observing = self.strategic_agent_ids
        else:
            rewarded = stage.rewarded_agents
            observing = self._stages[next_stage].acting_agents

        for aid in self.strategic_agent_ids:
            if self._done(aid):
                continue
            ctx = self._ctxs[aid]
            if aid in observing:
                obs = ctx.agent.encode_observation(ctx)
                if obs is not None:
                    observations[aid] = obs
                    infos[aid] = ctx.agent.collect_infos(ctx)
            if aid in rewarded:
                rewards[aid] = ctx.agent.compute_reward(ctx)
            self._flag_agent(aid, ctx, terminations, truncations)

        self._observations.update(observations)
        self._rewards.update(rewards)
        self._infos.update(infos)
        self.previous_stage, self._current_stage = self.current_stage, next_stage

        terminations["__all__"] = self.is_terminated()
        truncations["__all__"] = self.is_truncated()

        if self.current_stage is None or terminations["__all__"] or truncations["__all__"]:
            # terminal: flush the caches (fsm.py:360-375) -- the cache dicts themselves
            return self.Step(
                observations=self._observations,
                rewards=self._rewards,
                terminations=terminations,
                truncations=truncations,
                infos=self._infos,
            )
        # fsm.py:378: LAST computed reward of every agent that observes now (may be None)
        rewards = {aid: self._rewards[aid] for aid in observations}
        return self.Step(observations, rewards, terminations, truncations, infos)


# ========================================================================= Stackelberg
class StackelbergEnv(PhantomEnv):
    def __init__(
        self,
        num_steps: int,
        network: Network,
        leader_agents: Sequence[AgentID],
        follower_agents: Sequence[AgentID],
        env_supertype: Optional[Supertype] = None,
        agent_supertypes: Optional[Mapping[AgentID, Supertype]] = None,
    ):
        super().__init__(num_steps, network, env_supertype, agent_supertypes)
        for aid in list(leader_agents) + list(follower_agents):
            assert aid in network.agent_ids, f"Agent '{aid}' not in network"
        for aid in leader_agents:
            assert aid not in follower_agents, f"Agent '{aid}' not in network"
        self.leader_agents = leader_agents
        self.follower_agents = follower_agents
        self._rewards: Dict[AgentID, Optional[float]] = {}

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None):
        self._reset_common()
        self._rewards = {aid: None for aid in self.strategic_agent_ids}
        strategic = self.strategic_agent_ids
        return self._initial_obs([aid for aid in self.leader_agents if aid in strategic])

    def step(self, actions: Mapping[AgentID, Any]) -> PhantomEnv.Step:
        self._current_step += 1
        self._make_ctxs(self.agent_ids)
        if self.current_step % 2 == 1:  # stackelberg.py:133-137
            acting, observing = self.leader_agents, self.follower_agents
        else:
            acting, observing = self.follower_agents, self.leader_agents
        self._handle_acting_agents(acting, actions)
        self.resolve_network()

        observations, rewards, terminations, truncations, infos = {}, {}, {}, {}, {}
        for aid in self.strategic_agent_ids:
            if self._done(aid):
                continue
            ctx = self._ctxs[aid]
            if aid in observing:
                obs = ctx.agent.encode_observation(ctx)
                if obs is not None:
                    observations[aid] = obs
                    infos[aid] = ctx.agent.collect_infos(ctx)
            if aid in acting:
                self._rewards[aid] = ctx.agent.compute_reward(ctx)
            self._flag_agent(aid, ctx, terminations, truncations)

        terminations["__all__"] = self.is_terminated()
        truncations["__all__"] = self.is_truncated()
        if terminations["__all__"] or truncations["__all__"]:
            return self.Step(observations, self._rewards, terminations, truncations, infos)
        rewards = {
            aid: self._rewards[aid] for aid in observations if self._rewards[aid] is not None
        }
        return self.Step(observations, rewards, terminations, truncations, infos)
