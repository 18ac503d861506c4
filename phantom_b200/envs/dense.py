"""Dense-graph broadcast / batch-aggregation env on the device (family PHX_FAMILY_DENSE,
csrc/fam_dense.cu) -- BASELINE config C5.  Same env definition as oracle/workloads/dense.py
(which runs on the reference): N strategic agents on an adjacency-matrix graph, every agent
sends a receiver-tailored Signal to every neighbour, `handle_batch` aggregates (sum, max, first
arg-max) and Acks the arg-max sender; BatchResolver(round_limit=2)."""
from __future__ import annotations

import numpy as np

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box

N_AGENTS = 128


@ph.msg_payload("DenseAgent", "DenseAgent")
class Signal:
    value: int


@ph.msg_payload("DenseAgent", "DenseAgent")
class Ack:
    value: int


class DenseAgent(ph.StrategicAgent):
    """action [signal in 0..1]; obs [total/2^17, best/2^10, acks/2^7]; reward ack_total/2^10."""

    __phx_family__ = "dense"
    __phx_kind__ = 0
    __phx_device_class__ = True

    signal = device_column(0)
    total = device_column(1)
    best = device_column(2)
    best_sender = device_column(3)
    acks = device_column(4)
    ack_total = device_column(5)

    def __init__(self, agent_id):
        super().__init__(agent_id)
        self.observation_space = Box(0.0, 1.0, (3,))
        self.action_space = Box(0.0, 1.0, (1,))


FAMILY = register(FamilyInfo(
    name="dense",
    family_id=L.FAMILY_DENSE,
    payload_types=(Signal, Ack),
    obs_dim=3,
    act_dim=1,
    env_kinds=(L.ENV_BASE,),
    collect=lambda env, agents, spec: None,
    trace_capacity=lambda env, agents: len(agents) * len(agents),
))


class DenseEnv(ph.PhantomEnv):
    def __init__(self, n_agents: int = N_AGENTS, adjacency=None, *, num_steps: int = 8,
                 round_limit=2, enable_tracking: bool = False, **batch_kwargs):
        ids = [f"N{i}" for i in range(n_agents)]
        agents = [DenseAgent(a) for a in ids]
        network = ph.Network(agents, ph.resolvers.BatchResolver(
            enable_tracking=enable_tracking, round_limit=round_limit))
        if adjacency is None:
            adjacency = np.ones((n_agents, n_agents), np.int64) - np.eye(n_agents, dtype=np.int64)
        network.add_connections_with_adjmat(ids, np.asarray(adjacency))
        self.ids = ids
        super().__init__(num_steps=num_steps, network=network, **batch_kwargs)
