"""Leader-follower pricing game on the device (family PHX_FAMILY_STACKELBERG,
csrc/fam_stackelberg.cu) -- BASELINE config C4.  Same env definition as
oracle/workloads/stackelberg.py (which runs on the reference)."""
from __future__ import annotations

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box

N_FOLLOWERS = 3
CAPACITY = 15
KIND_LEADER, KIND_FOLLOWER = 0, 1


@ph.msg_payload("LeaderAgent", "FollowerAgent")
class Price:
    ticks: int


@ph.msg_payload("FollowerAgent", "LeaderAgent")
class Demand:
    qty: int


@ph.msg_payload("LeaderAgent", "FollowerAgent")
class Ack:
    filled: int


class LeaderAgent(ph.StrategicAgent):
    """action [price in 0..1]; serves demands first come first served from CAPACITY units;
    obs [demand/30, remaining/CAPACITY]; reward = revenue of the round / 100."""

    __phx_family__ = "stackelberg_game"
    __phx_kind__ = KIND_LEADER
    __phx_device_class__ = True

    price = device_column(0)
    remaining = device_column(1)
    revenue_round = device_column(2)
    demand_round = device_column(3)

    def __init__(self, agent_id):
        super().__init__(agent_id)
        self.observation_space = Box(0.0, 1.0, (2,))
        self.action_space = Box(0.0, 1.0, (1,))


class FollowerAgent(ph.StrategicAgent):
    """action [quantity in 0..1 -> 0..10 units]; obs [price/100, filled/10];
    reward = filled * (value - price) / 100."""

    __phx_family__ = "stackelberg_game"
    __phx_kind__ = KIND_FOLLOWER
    __phx_device_class__ = True

    value = device_column(0)
    seen_price = device_column(1)
    last_filled = device_column(2)
    utility_round = device_column(3)

    def __init__(self, agent_id, leader_id):
        super().__init__(agent_id)
        self.leader_id = leader_id
        self.observation_space = Box(0.0, 1.0, (2,))
        self.action_space = Box(0.0, 1.0, (1,))


def _collect(env, agents, spec) -> None:
    leaders = [a for a in agents if isinstance(a, LeaderAgent)]
    if len(leaders) != 1:
        raise NotLowerableError("stackelberg device program: exactly one LeaderAgent")
    spec.iparams[0] = int(getattr(env, "capacity", CAPACITY))
    k = 0
    for a in agents:
        if isinstance(a, FollowerAgent):
            if a.leader_id != leaders[0].id:
                raise NotLowerableError(f"follower '{a.id}' addresses unknown leader")
            spec.agent_iparam[a._phx_slot][0] = k
            spec.agent_iparam[a._phx_slot][1] = leaders[0]._phx_slot
            k += 1


FAMILY = register(FamilyInfo(
    name="stackelberg_game",
    family_id=L.FAMILY_STACKELBERG,
    payload_types=(Price, Demand, Ack),
    obs_dim=2,
    act_dim=1,
    env_kinds=(L.ENV_STACKELBERG,),
    collect=_collect,
    trace_capacity=lambda env, agents: 4 * len(agents),
))


class StackelbergGameEnv(ph.StackelbergEnv):
    def __init__(self, n_followers: int = N_FOLLOWERS, *, num_steps: int = 100,
                 enable_tracking: bool = False, **batch_kwargs):
        follower_ids = [f"F{i + 1}" for i in range(n_followers)]
        agents = [LeaderAgent("LEADER")] + [FollowerAgent(f, "LEADER") for f in follower_ids]
        network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking))
        network.add_connections_between(["LEADER"], follower_ids)
        self.follower_ids = follower_ids
        super().__init__(num_steps, network, ["LEADER"], follower_ids, **batch_kwargs)
