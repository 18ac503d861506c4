"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Network + resolver of the reference, restated.

Reference lines followed (all under /root/reference/phantom/):
  network.py:31        NetworkError
  network.py:59-85     Network.__init__
  network.py:87-177    add_agent(s) / add_connection* / adjacency-matrix validation
  network.py:179-184   reset
  network.py:186-206   subnet_for
  network.py:208-222   context_for
  network.py:224-231   has_edge
  network.py:233-254   send
  network.py:256-265   resolve
  network.py:267-295   get_agents_*
  network.py:297-331   _enforce_payload_checks
  network.py:340-453   StochasticNetwork
  resolvers.py:17-88   Resolver
  resolvers.py:91-163  BatchResolver

The reference stores the graph in a networkx.DiGraph; the only properties the hot path
relies on are (a) edge membership and (b) neighbour iteration in edge-insertion order
(context_for builds agent_views in that order).  A dict-of-dicts keeps both.
"""
from __future__ import annotations

import itertools
import warnings
from abc import ABC, abstractmethod
from collections import defaultdict
from copy import deepcopy
from typing import Callable, Dict, Iterable, List, Mapping, Optional, Sequence, Tuple

import numpy as np

from .core import Agent, AgentID, Context, EnvView, Message, MsgPayload


class NetworkError(Exception):
    pass


class _DiGraph:
    """Insertion-ordered directed graph with the few networkx calls the reference uses."""

    def __init__(self):
        self._succ: Dict[AgentID, Dict[AgentID, None]] = {}
        self._pred: Dict[AgentID, Dict[AgentID, None]] = {}

    def add_node(self, n):
        self._succ.setdefault(n, {})
        self._pred.setdefault(n, {})

    def add_edge(self, u, v):
        self.add_node(u)
        self.add_node(v)
        self._succ[u][v] = None
        self._pred[v][u] = None

    def has_edge(self, u, v) -> bool:
        return u in self._succ and v in self._succ[u]

    def neighbors(self, n):
        return iter(self._succ[n])

    successors = neighbors

    def predecessors(self, n):
        return iter(self._pred[n])

    @property
    def nodes(self):
        return self._succ.keys()

    @property
    def edges(self):
        return _EdgeView(self)

    def subgraph(self, nodes):
        keep = list(dict.fromkeys(nodes))
        g = _DiGraph()
        for n in self._succ:  # networkx subgraph views keep the parent's node order
            if n in keep:
                g.add_node(n)
        for u in g._succ:
            for v in self._succ[u]:
                if v in g._succ:
                    g.add_edge(u, v)
        return g

    def __len__(self):
        return len(self._succ)


class _EdgeView:
    def __init__(self, g):
        self._g = g

    def __contains__(self, uv):
        return self._g.has_edge(*uv)

    def __iter__(self):
        for u, vs in self._g._succ.items():
            for v in vs:
                yield (u, v)

    def __len__(self):
        return sum(len(vs) for vs in self._g._succ.values())


# ---------------------------------------------------------------------------- resolver
class Resolver(ABC):
    def __init__(self, enable_tracking: bool = False):
        self.enable_tracking = enable_tracking
        self._tracked_messages: List[Message] = []

    def push(self, message: Message) -> None:
        if self.enable_tracking:  # resolvers.py:41-42
            self._tracked_messages.append(message)
        self.handle_push(message)

    def clear_tracked_messages(self) -> None:
        self._tracked_messages.clear()

    @property
    def tracked_messages(self) -> List[Message]:
        return self._tracked_messages

    @abstractmethod
    def handle_push(self, message: Message) -> None:
        raise NotImplementedError

    @abstractmethod
    def resolve(self, network: "Network", contexts: Mapping[AgentID, Context]) -> None:
        raise NotImplementedError

    @abstractmethod
    def reset(self) -> None:
        raise NotImplementedError


class BatchResolver(Resolver):
    def __init__(
        self,
        enable_tracking: bool = False,
        round_limit: Optional[int] = None,
        shuffle_batches: bool = False,
    ):
        super().__init__(enable_tracking)
        self.round_limit = round_limit
        self.shuffle_batches = shuffle_batches
        # receiver -> batch; dict insertion order == receiver FIRST-ARRIVAL order
        self.messages: Dict[AgentID, List[Message]] = defaultdict(list)

    def reset(self) -> None:
        self.messages.clear()

    def handle_push(self, message: Message) -> None:
        self.messages[message.receiver_id].append(message)  # resolvers.py:125-126

    def resolve(self, network: "Network", contexts: Mapping[AgentID, Context]) -> None:
        rounds = (
            itertools.count() if self.round_limit is None else range(self.round_limit)
        )
        for _ in rounds:  # resolvers.py:133-158
            if not self.messages:
                break
            inbox, self.messages = self.messages, defaultdict(list)
            for receiver_id, batch in inbox.items():
                if receiver_id not in contexts:
                    continue  # done agent: mail dropped silently (resolvers.py:143-144)
                live = [m for m in batch if network.has_edge(m.sender_id, m.receiver_id)]
                if self.shuffle_batches:
                    np.random.shuffle(live)
                ctx = contexts[receiver_id]
                responses = ctx.agent.handle_batch(ctx, live)
                if responses is not None:
                    for sub_receiver, sub_payload in responses:
                        network.send(receiver_id, sub_receiver, sub_payload)
        if self.messages:  # resolvers.py:160-163
            raise RuntimeError(
                f"{len(self.messages)} message(s) still in queue after BatchResolver "
                "round limit reached."
            )


# ----------------------------------------------------------------------------- network
class Network:
    def __init__(
        self,
        agents: Optional[Iterable[Agent]] = None,
        resolver: Optional[Resolver] = None,
        connections: Optional[Iterable[Tuple[AgentID, AgentID]]] = None,
        ignore_connection_errors: bool = False,
        enforce_msg_payload_checks: bool = True,
    ):
        self.graph = _DiGraph()
        self.agents: Dict[AgentID, Agent] = {}
        self.resolver = resolver or BatchResolver()  # round_limit=None (network.py:69)
        self.ignore_connection_errors = ignore_connection_errors
        self.enforce_msg_payload_checks = enforce_msg_payload_checks
        self._warned_deprecated_payload = False
        if agents is not None:
            self.add_agents(agents)
        if connections is not None:
            for c in connections:
                self.add_connection(*c)

    @property
    def agent_ids(self):
        return self.agents.keys()

    def add_agent(self, agent: Agent) -> None:
        if agent.id in self.agents:
            raise ValueError(f"Agent with ID = '{agent.id}' already exists.")
        self.agents[agent.id] = agent
        self.graph.add_node(agent.id)

    def add_agents(self, agents: Iterable[Agent]) -> None:
        for a in agents:
            self.add_agent(a)

    def add_connection(self, u: AgentID, v: AgentID) -> None:
        for x in (u, v):
            if x not in self.agents:
                raise ValueError(f"Agent with ID = '{x}' does not exist.")
        self.graph.add_edge(u, v)  # always both directions (network.py:122-123)
        self.graph.add_edge(v, u)

    def add_connections_from(self, ebunch) -> None:
        for u, v in ebunch:
            self.add_connection(u, v)

    def add_connections_between(self, us, vs) -> None:
        self.add_connections_from(itertools.product(us, vs))

    def add_connections_with_adjmat(self, agent_ids: Sequence[AgentID], adjacency_matrix) -> None:
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
                [aid], [agent_ids[j] for j in range(n) if adjacency_matrix[i, j] > 0]
            )

    def reset(self) -> None:
        self.resolver.reset()
        for agent in self.agents.values():
            agent.reset()

    def subnet_for(self, agent_id: AgentID) -> "Network":
        sub = Network.__new__(Network)
        sub.graph = self.graph.subgraph(
            itertools.chain(
                (agent_id,),
                self.graph.successors(agent_id),
                self.graph.predecessors(agent_id),
            )
        )
        sub.agents = {aid: self.agents[aid] for aid in sub.graph.nodes}
        sub.resolver = deepcopy(self.resolver)
        sub.resolver.reset()
        return sub

    def context_for(self, agent_id: AgentID, env_view: EnvView) -> Context:
        views = {
            nid: self.agents[nid].view(agent_id) for nid in self.graph.neighbors(agent_id)
        }
        return Context(self.agents[agent_id], views, env_view)

    def has_edge(self, sender_id: AgentID, receiver_id: AgentID) -> bool:
        return self.graph.has_edge(sender_id, receiver_id)

    def send(self, sender_id: AgentID, receiver_id: AgentID, payload) -> None:
        # network.py:246-254: edge check, then whitelist check, then push.
        if not self.ignore_connection_errors and not self.has_edge(sender_id, receiver_id):
            raise NetworkError(f"No connection between {sender_id} and {receiver_id}.")
        if self.enforce_msg_payload_checks:
            self._enforce_payload_checks(sender_id, receiver_id, payload)
        self.resolver.push(Message(sender_id, receiver_id, payload))

    def resolve(self, contexts: Mapping[AgentID, Context]) -> None:
        self.resolver.resolve(self, contexts)
        self.resolver.reset()

    def get_agents_where(self, pred: Callable[[Agent], bool]) -> Dict[AgentID, Agent]:
        return {aid: self.agents[aid] for aid in self.graph.nodes if pred(self.agents[aid])}

    def get_agents_with_type(self, agent_type) -> Dict[AgentID, Agent]:
        return self.get_agents_where(lambda a: isinstance(a, agent_type))

    def get_agents_without_type(self, agent_type) -> Dict[AgentID, Agent]:
        return self.get_agents_where(lambda a: not isinstance(a, agent_type))

    def _enforce_payload_checks(self, sender_id, receiver_id, payload) -> None:
        if not hasattr(payload, "_sender_types") or not hasattr(payload, "_receiver_types"):
            if isinstance(payload, MsgPayload):  # deprecated base class: warn once
                if not self._warned_deprecated_payload:
                    warnings.warn(
                        "MsgPayload type is deprecated. In future, use the @msg_payload decorator",
                        DeprecationWarning,
                    )
                    self._warned_deprecated_payload = True
                return
            raise NetworkError(
                "Message payloads sent across the network must use the 'msg_payload' "
                f"decorator (bad payload = '{payload}')"
            )
        sender, receiver = self.agents[sender_id], self.agents[receiver_id]
        # exact class *name* match (network.py:315-331)
        if payload._sender_types is not None and type(sender).__name__ not in payload._sender_types:
            raise NetworkError(
                f"Message payload of type '{type(payload).__name__}' cannot be sent by "
                f"agent with type '{type(sender).__name__}' (expected one of {payload._sender_types})"
            )
        if (
            payload._receiver_types is not None
            and type(receiver).__name__ not in payload._receiver_types
        ):
            raise NetworkError(
                f"Message payload of type '{type(payload).__name__}' cannot be received by "
                f"agent with type '{type(receiver).__name__}' (expected one of {payload._receiver_types})"
            )

    def __getitem__(self, agent_id: AgentID) -> Agent:
        return self.agents[agent_id]

    def __len__(self) -> int:
        return len(self.graph)


class StochasticNetwork(Network):
    """network.py:340-453: every connection carries a rate; edges are re-drawn with
    np.random.random() at construction and on every reset()."""

    def __init__(self, agents=None, resolver=None, connections=None,
                 ignore_connection_errors=False, enforce_msg_payload_checks=True):
        # the reference creates _base_connections only *after* super().__init__
        # (network.py:370-378), so `connections=` given to this constructor fails with
        # AttributeError inside add_connection; kept as is.
        super().__init__(agents, resolver, connections, ignore_connection_errors,
                         enforce_msg_payload_checks)
        self._base_connections: List[Tuple[AgentID, AgentID, float]] = []

    def add_connection(self, u: AgentID, v: AgentID, rate: float = 1.0) -> None:
        if np.random.random() < rate:
            self.graph.add_edge(u, v)
            self.graph.add_edge(v, u)
        self._base_connections.append((u, v, rate))

    def add_connections_from(self, ebunch) -> None:
        for c in ebunch:
            if len(c) == 2:
                self.add_connection(c[0], c[1])
            elif len(c) == 3:
                self.add_connection(c[0], c[1], c[2])
            else:
                raise ValueError(f"Ill-formatted connection tuple {c}.")

    def add_connections_between(self, us, vs, rate: float = 1.0) -> None:
        for u, v in itertools.product(us, vs):
            self.add_connection(u, v, rate)

    def resample_connectivity(self) -> None:
        self.graph = _DiGraph()
        for agent in self.agents.values():
            self.graph.add_node(agent.id)
        for u, v, rate in self._base_connections:
            if np.random.random() < rate:
                self.graph.add_edge(u, v)
                self.graph.add_edge(v, u)

    def reset(self) -> None:
        self.resample_connectivity()
        Network.reset(self)
