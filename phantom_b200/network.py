"""Network (reference: phantom/network.py:35-337).

Host side: the agent table (insertion ordered == slot order on the device) and the directed
graph, with the reference's construction API and validation errors.  `send` / `resolve`
are device code.  Lowering turns the graph into per-slot adjacency bitmasks.
"""
from __future__ import annotations

import itertools
from typing import Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from .agents import Agent
from .context import Context
from .errors import DeviceOnlyError
from .resolvers import BatchResolver, Resolver
from .types import AgentID
from .views import EnvView


class NetworkError(Exception):
    pass


class Network:
    def __init__(self, agents: Optional[Iterable[Agent]] = None,
                 resolver: Optional[Resolver] = None,
                 connections: Optional[Iterable[Tuple[AgentID, AgentID]]] = None,
                 ignore_connection_errors: bool = False,
                 enforce_msg_payload_checks: bool = True) -> None:
        self.agents: Dict[AgentID, Agent] = {}
        self._succ: Dict[AgentID, Dict[AgentID, None]] = {}
        self.resolver = resolver or BatchResolver()
        self.ignore_connection_errors = ignore_connection_errors
        self.enforce_msg_payload_checks = enforce_msg_payload_checks
        if agents is not None:
            self.add_agents(agents)
        if connections is not None:
            for c in connections:
                self.add_connection(*c)

    # ------------------------------------------------------------------ construction
    @property
    def agent_ids(self):
        return self.agents.keys()

    def add_agent(self, agent: Agent) -> None:
        if agent.id in self.agents:
            raise ValueError(f"Agent with ID = '{agent.id}' already exists.")
        self.agents[agent.id] = agent
        self._succ[agent.id] = {}

    def add_agents(self, agents: Iterable[Agent]) -> None:
        for a in agents:
            self.add_agent(a)

    def add_connection(self, u: AgentID, v: AgentID) -> None:
        for x in (u, v):
            if x not in self.agents:
                raise ValueError(f"Agent with ID = '{x}' does not exist.")
        self._succ[u][v] = None
        self._succ[v][u] = None

    def add_connections_from(self, ebunch: Iterable[Tuple[AgentID, AgentID]]) -> None:
        for u, v in ebunch:
            self.add_connection(u, v)

    def add_connections_between(self, us: Iterable[AgentID], vs: Iterable[AgentID]) -> None:
        self.add_connections_from(itertools.product(us, vs))

    def add_connections_with_adjmat(self, agent_ids: Sequence[AgentID],
                                    adjacency_matrix: np.ndarray) -> None:
        n = adjacency_matrix.shape[0]
        if len(agent_ids) != n:
            raise ValueError("Number of agent IDs doesn't match adjacency matrix dimensions.")
        if len(set(adjacency_matrix.shape)) != 1:
            raise ValueError("Adjacency matrix must be square.")
        if not (adjacency_matrix.transpose() == adjacency_matrix).all():
            raise ValueError("Adjacency matrix must be symmetric.")
        if not (np.abs(adjacency_matrix.diagonal()) < 1e-5).all():
            raise ValueError("Adjacency matrix must be hollow.")
        for i, aid in enumerate(agent_ids):
            self.add_connections_between(
                [aid], [agent_ids[j] for j in range(n) if adjacency_matrix[i, j] > 0])

    # ------------------------------------------------------------------------ queries
    def has_edge(self, sender_id: AgentID, receiver_id: AgentID) -> bool:
        return sender_id in self._succ and receiver_id in self._succ[sender_id]

    def neighbours(self, agent_id: AgentID) -> List[AgentID]:
        return list(self._succ[agent_id])

    def context_for(self, agent_id: AgentID, env_view: EnvView) -> Context:
        views = {n: self.agents[n].view(agent_id) for n in self._succ[agent_id]}
        return Context(self.agents[agent_id], views, env_view)

    def subnet_for(self, agent_id: AgentID) -> "Network":
        keep = {agent_id, *self._succ[agent_id],
                *(u for u, vs in self._succ.items() if agent_id in vs)}
        sub = Network.__new__(Network)
        sub.agents = {aid: a for aid, a in self.agents.items() if aid in keep}
        sub._succ = {u: {v: None for v in vs if v in keep}
                     for u, vs in self._succ.items() if u in keep}
        sub.resolver = type(self.resolver).__new__(type(self.resolver))
        sub.resolver.__dict__.update(self.resolver.__dict__)
        sub.resolver._tracked_messages = []
        sub.ignore_connection_errors = self.ignore_connection_errors
        sub.enforce_msg_payload_checks = self.enforce_msg_payload_checks
        return sub

    def get_agents_where(self, pred: Callable[[Agent], bool]) -> Dict[AgentID, Agent]:
        return {aid: a for aid, a in self.agents.items() if pred(a)}

    def get_agents_with_type(self, agent_type) -> Dict[AgentID, Agent]:
        return self.get_agents_where(lambda a: isinstance(a, agent_type))

    def get_agents_without_type(self, agent_type) -> Dict[AgentID, Agent]:
        return self.get_agents_where(lambda a: not isinstance(a, agent_type))

    def adjacency_matrix(self) -> np.ndarray:
        ids = list(self.agents)
        pos = {a: i for i, a in enumerate(ids)}
        m = np.zeros((len(ids), len(ids)), np.uint8)
        for u, vs in self._succ.items():
            for v in vs:
                m[pos[u], pos[v]] = 1
        return m

    # ------------------------------------------------------------- device-side methods
    def reset(self) -> None:
        self.resolver.reset()
        for agent in self.agents.values():
            agent.reset()

    def send(self, sender_id, receiver_id, payload) -> None:
        raise DeviceOnlyError("Network.send runs inside the fused step kernel")

    def resolve(self, contexts: Mapping[AgentID, Context]) -> None:
        raise DeviceOnlyError("Network.resolve runs inside the fused step kernel")

    def __getitem__(self, agent_id: AgentID) -> Agent:
        return self.agents[agent_id]

    def __len__(self) -> int:
        return len(self.agents)


class StochasticNetwork(Network):
    """Network whose edges are re-drawn at every reset (reference: phantom/network.py:340-453).

    Host side keeps `_base_connections` = [(u, v, rate)] in insertion order; the graph the
    host object shows (`has_edge`, `neighbours`, `adjacency_matrix`) is the set of POSSIBLE
    edges (rate > 0).  The per-env, per-episode graphs live on the device: every env resamples
    base connection c at reset with `uniform01(stream 5, step 0, idx c) < rate`
    (csrc/phx_engine.cuh `resample_adj_row`); read them with `PhantomEnv.adjacency()`.

    As in the reference (network.py:370-378) the constructor does not accept `connections`
    (its `_base_connections` list does not exist yet when the base constructor runs).
    """

    def __init__(self, agents: Optional[Iterable[Agent]] = None,
                 resolver: Optional[Resolver] = None,
                 connections: Optional[Iterable[Tuple[AgentID, AgentID]]] = None,
                 ignore_connection_errors: bool = False,
                 enforce_msg_payload_checks: bool = True) -> None:
        super().__init__(agents, resolver, connections, ignore_connection_errors,
                         enforce_msg_payload_checks)
        self._base_connections: List[Tuple[AgentID, AgentID, float]] = []

    def add_connection(self, u: AgentID, v: AgentID, rate: float = 1.0) -> None:
        for x in (u, v):
            if x not in self.agents:
                raise ValueError(f"Agent with ID = '{x}' does not exist.")
        rate = float(rate)
        if rate > 0.0:
            self._succ[u][v] = None
            self._succ[v][u] = None
        self._base_connections.append((u, v, rate))

    def add_connections_from(self, ebunch) -> None:
        for c in ebunch:
            if len(c) == 2:
                self.add_connection(c[0], c[1])
            elif len(c) == 3:
                self.add_connection(c[0], c[1], c[2])
            else:
                raise ValueError(f"Ill-formatted connection tuple {c}.")

    def add_connections_between(self, us: Iterable[AgentID], vs: Iterable[AgentID],
                                rate: float = 1.0) -> None:
        for u, v in itertools.product(us, vs):
            self.add_connection(u, v, rate)

    def resample_connectivity(self) -> None:
        raise DeviceOnlyError("connectivity is resampled per env by the reset kernel")
