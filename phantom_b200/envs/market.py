"""Three-stage FSM market on the device (family PHX_FAMILY_MARKET, csrc/fam_market.cu) --
BASELINE config C3.  Same env definition as oracle/workloads/market.py (which runs on the
reference): makers quote, takers order through a clearing agent with first-come-first-served
capacity, the clearing agent settles; MAKER -> TAKER -> CLEARING -> MAKER.
"""
from __future__ import annotations

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box, Discrete

N_MAKERS, N_TAKERS = 7, 24
MAKER_INVENTORY = 12
MAKER_CAPACITY = 3
KIND_MAKER, KIND_TAKER, KIND_CLEARING = 0, 1, 2


@ph.msg_payload("MakerAgent", "TakerAgent")
class Quote:
    price: int


@ph.msg_payload("TakerAgent", "ClearingAgent")
class Order:
    maker: int
    price: int


@ph.msg_payload("ClearingAgent", ["MakerAgent", "TakerAgent"])
class Fill:
    units: int
    notional: int


class MakerAgent(ph.StrategicAgent):
    """action [price in 0..1] -> Quote(rint(100 a)) to every taker; obs [inventory/INV,
    last_price/100]; reward = last notional / 100; terminates when sold out."""

    __phx_family__ = "market"
    __phx_kind__ = KIND_MAKER
    __phx_device_class__ = True

    inventory = device_column(0)
    cash = device_column(1)
    last_price = device_column(2)
    last_notional = device_column(3)

    def __init__(self, agent_id):
        super().__init__(agent_id)
        self.observation_space = Box(0.0, 1.0, (2,))
        self.action_space = Box(0.0, 1.0, (1,))


class TakerAgent(ph.StrategicAgent):
    """keeps the best quote of the cycle; action Discrete(2): 1 = buy at the best quote;
    obs [value/100, quote/100, holdings/33]; reward = surplus of the cycle / 100."""

    __phx_family__ = "market"
    __phx_kind__ = KIND_TAKER
    __phx_device_class__ = True

    value = device_column(0)
    best_price = device_column(1)
    best_maker = device_column(2)
    holdings = device_column(3)
    last_surplus = device_column(4)

    def __init__(self, agent_id):
        super().__init__(agent_id)
        self.observation_space = Box(0.0, 1.0, (3,))
        self.action_space = Discrete(2)


class ClearingAgent(ph.Agent):
    """accepts orders first come first served up to MAKER_CAPACITY per maker and cycle, then
    settles with Fill messages in the CLEARING stage."""

    __phx_family__ = "market"
    __phx_kind__ = KIND_CLEARING
    __phx_device_class__ = True


def _collect(env, agents, spec) -> None:
    makers = [a for a in agents if isinstance(a, MakerAgent)]
    takers = [a for a in agents if isinstance(a, TakerAgent)]
    clearing = [a for a in agents if isinstance(a, ClearingAgent)]
    if len(clearing) != 1:
        raise NotLowerableError("market device program: exactly one ClearingAgent")
    stage_ids = list(env._stages)
    for needed in ("MAKER", "CLEARING"):
        if needed not in stage_ids:
            raise NotLowerableError(f"market device program needs a stage called '{needed}'")
    spec.iparams[0], spec.iparams[1] = len(makers), len(takers)
    spec.iparams[2] = int(getattr(env, "maker_inventory", MAKER_INVENTORY))
    spec.iparams[3] = int(getattr(env, "maker_capacity", MAKER_CAPACITY))
    spec.iparams[4] = stage_ids.index("MAKER")
    spec.iparams[5] = stage_ids.index("CLEARING")
    spec.iparams[6] = clearing[0]._phx_slot
    for group in (makers, takers):
        for k, a in enumerate(group):
            spec.agent_iparam[a._phx_slot][0] = k


FAMILY = register(FamilyInfo(
    name="market",
    family_id=L.FAMILY_MARKET,
    payload_types=(Quote, Order, Fill),
    obs_dim=3,
    act_dim=1,
    env_kinds=(L.ENV_FSM,),
    collect=_collect,
    trace_capacity=lambda env, agents: 8 * len(agents),
))


class MarketEnv(ph.FiniteStateMachineEnv):
    def __init__(self, n_makers: int = N_MAKERS, n_takers: int = N_TAKERS, *, num_steps: int = 99,
                 enable_tracking: bool = False, shuffle_batches: bool = False, **batch_kwargs):
        maker_ids = [f"M{i + 1}" for i in range(n_makers)]
        taker_ids = [f"T{i + 1}" for i in range(n_takers)]
        agents = [MakerAgent(m) for m in maker_ids] + [TakerAgent(t) for t in taker_ids]
        agents.append(ClearingAgent("CLEARING"))
        network = ph.Network(agents, ph.resolvers.BatchResolver(
            enable_tracking=enable_tracking, shuffle_batches=shuffle_batches))
        network.add_connections_between(maker_ids, taker_ids)
        network.add_connections_between(["CLEARING"], maker_ids + taker_ids)
        everyone = maker_ids + taker_ids
        self.maker_ids, self.taker_ids = maker_ids, taker_ids
        super().__init__(
            num_steps=num_steps, network=network, initial_stage="MAKER",
            stages=[
                ph.FSMStage("MAKER", acting_agents=maker_ids, rewarded_agents=[],
                            next_stages=["TAKER"]),
                ph.FSMStage("TAKER", acting_agents=taker_ids, rewarded_agents=[],
                            next_stages=["CLEARING"]),
                ph.FSMStage("CLEARING", acting_agents=["CLEARING"], rewarded_agents=everyone,
                            next_stages=["MAKER"]),
            ], **batch_kwargs)
