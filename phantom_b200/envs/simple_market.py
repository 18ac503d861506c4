"""The reference's second example environment on the device (family PHX_FAMILY_SIMPLE_MARKET,
csrc/fam_simple_market.cu): examples/environments/simple_market/ -- buyers and sellers in a
two-stage FiniteStateMachineEnv whose ENV CLASS keeps state of its own (`avg_price`), publishes
it through a custom EnvView field and updates it in an env-level `post_message_resolution`
(simple_mkt_env.py:9-58, market_agents.py:33-132).  Same class names, constructor signatures and
stage layout as the example, so example_simple_market.py:9-30 builds this env unchanged.
"""
from __future__ import annotations

import dataclasses
import math

import numpy as np

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.fsm import FSMEnvView
from phantom_b200.spaces import Box, Discrete
from phantom_b200.utils.samplers import KIND_UNIFORM_FLOAT, Sampler

KIND_BUYER, KIND_SELLER = 0, 1
MAX_SELLERS = 15


def buyer_layout(n_sellers: int):
    """State-word layout of a buyer (csrc/fam_simple_market.cu SimpleMarketProgramT<MAXS>): the
    device program is built for MAXS = 7 remembered sellers (19 words) or 15 (36 words).
    -> (word of the dict order, words it takes, word of current_reward, word of type.value)"""
    maxs = 7 if n_sellers <= 7 else 15
    w_ord, ow = 2 * maxs, (1 if maxs == 7 else 2)
    return w_ord, ow, w_ord + ow, w_ord + ow + 2


@ph.msg_payload()
class Price:
    price: float

    @classmethod
    def _phx_decode(cls, lo: int, hi: int) -> "Price":
        """the device payload is the float64 price, low / high word"""
        bits = (np.uint64(hi & 0xFFFFFFFF) << np.uint64(32)) | np.uint64(lo & 0xFFFFFFFF)
        return cls(float(np.array([bits], np.uint64).view(np.float64)[0]))


@ph.msg_payload()
class Order:
    vol: int


@dataclasses.dataclass
class BuyerSupertype(ph.Supertype):
    """Buyer type = the buyer's intrinsic value for the good (market_agents.py:27-30)."""

    value: float


def _f64_column(env, agent, word):
    lo = env.agent_column(agent, word).astype(np.uint32).astype(np.uint64)
    hi = env.agent_column(agent, word + 1).astype(np.uint32).astype(np.uint64)
    v = ((hi << np.uint64(32)) | lo).view(np.float64)
    return v.item() if v.size == 1 else v


class BuyerAgent(ph.StrategicAgent):
    """action Discrete(2): 1 = buy one unit from a cheapest seller; obs [min price, demand,
    type.value]; reward = sum over the cycle of (value - price paid)."""

    __phx_family__ = "simple_market"
    __phx_kind__ = KIND_BUYER
    __phx_device_class__ = True

    def __init__(self, agent_id, demand_prob, supertype):
        super().__init__(agent_id, supertype=supertype)
        self.demand_prob = demand_prob
        self.action_space = Discrete(2)
        self.observation_space = Box(low=0, high=1, shape=(3,))

    def _layout(self):
        return buyer_layout(int(self._phx_env.spec.iparams[0]))

    @property
    def current_reward(self):
        return _f64_column(self._phx_env, self, self._layout()[2])

    @property
    def type(self):
        return BuyerSupertype(value=_f64_column(self._phx_env, self, self._layout()[3]))

    @property
    def n_sellers_heard(self):
        """len(self.seller_prices): the top nibble of the dict-order words."""
        w_ord, ow, _, _ = self._layout()
        return (self._phx_env.agent_column(self, w_ord + ow - 1).astype(np.uint32) >> 28).astype(np.int64)


class SellerAgent(ph.StrategicAgent):
    """action Box(0, 1): the price, sent to every neighbour; obs [volume transacted,
    env_view.avg_price]; reward = revenue of the cycle.  Infinite supply."""

    __phx_family__ = "simple_market"
    __phx_kind__ = KIND_SELLER
    __phx_device_class__ = True

    def __init__(self, agent_id):
        super().__init__(agent_id)
        self.action_space = Box(low=0, high=1, shape=(1,))
        self.observation_space = Box(np.array([0, 0]), np.array([np.inf, 1]))

    @property
    def current_price(self):
        return _f64_column(self._phx_env, self, 0)

    @property
    def current_revenue(self):
        return _f64_column(self._phx_env, self, 2)

    @property
    def current_tx(self):
        return _f64_column(self._phx_env, self, 4)


def _collect(env, agents, spec) -> None:
    buyers = [a for a in agents if isinstance(a, BuyerAgent)]
    sellers = [a for a in agents if isinstance(a, SellerAgent)]
    if not buyers or not 1 <= len(sellers) <= MAX_SELLERS:
        raise NotLowerableError("simple-market device program: >= 1 buyer and 1..15 sellers")
    if not isinstance(env, SimpleMarketEnv):
        raise NotLowerableError("simple-market agents run under SimpleMarketEnv")
    first = env._stages[env.initial_stage].acting_agents
    if {a.id for a in sellers} - set(first):
        raise NotLowerableError(
            "simple-market device program: the sellers must act in the initial stage (a buyer "
            "that observes before any price was heard raises ValueError in the reference)")
    # SellerAgent.decode_action prices `ctx.neighbour_ids` in the graph's own (insertion) order
    # (market_agents.py:108-112), which is the push order of the Price messages; the device program
    # walks a seller's neighbours in slot order, so the two must agree
    for a in sellers:
        order = [env.agents[n]._phx_slot for n in env.network.neighbours(a.id)]
        if order != sorted(order):
            raise NotLowerableError(
                f"simple-market device program: the neighbours of seller '{a.id}' were connected in "
                "an order that differs from the network's agent order; the reference would price "
                "them in connection order.  Add the connections in agent order (e.g. "
                "add_connections_between(buyer_ids, seller_ids) with buyer_ids in network order)")
    spec.iparams[0] = len(sellers)
    draw = 0
    for group in (buyers, sellers):
        for k, a in enumerate(group):
            spec.agent_iparam[a._phx_slot][0] = k
    for a in agents:  # reset-time draws happen in agent order (Network.reset, network.py:179-184)
        i = a._phx_slot
        if isinstance(a, SellerAgent):
            low, high = float(a.action_space.low.reshape(-1)[0]), float(a.action_space.high.reshape(-1)[0])
            spec.agent_iparam[i][1] = draw
            spec.agent_fparam[i][0], spec.agent_fparam[i][1] = low, high
            draw += 1
            continue
        if a.supertype is None:
            raise NotLowerableError(f"buyer '{a.id}' needs a BuyerSupertype")
        v = a.supertype.value
        if isinstance(v, Sampler):
            if getattr(a.supertype, "_managed", False):
                raise NotLowerableError("simple-market: pass supertypes to the agents, as the example does")
            kind, low, high = v.device_desc()
            if kind != KIND_UNIFORM_FLOAT:
                raise NotLowerableError("BuyerSupertype.value: only UniformFloatSampler is lowered")
            spec.agent_iparam[i][1] = draw
            spec.agent_fparam[i][0], spec.agent_fparam[i][1] = low, high
            draw += 1
        else:
            spec.agent_iparam[i][1] = -1
            spec.agent_fparam[i][0] = float(v)
        p = float(a.demand_prob)
        if not 0.0 <= p <= 1.0:
            raise NotLowerableError("demand_prob must be a probability")
        spec.agent_iparam[i][3] = int(math.ceil(p * 16777216.0))  # uniform01 < p <=> d24 < ceil(p 2^24)


FAMILY = register(FamilyInfo(
    name="simple_market",
    family_id=L.FAMILY_SIMPLE_MARKET,
    payload_types=(Price, Order),
    obs_dim=3,
    act_dim=1,
    env_kinds=(L.ENV_FSM,),
    collect=_collect,
    trace_capacity=lambda env, agents: len(agents) * len(agents),
    supports_supertypes=True,
))


class SimpleMarketEnv(ph.FiniteStateMachineEnv):
    """simple_mkt_env.py:9-58.  `avg_price` lives on the device (env-level words 0/1, float64),
    is updated by the kernel's env-level post hook and reaches the sellers' observations through
    the start-of-step EnvView snapshot."""

    __phx_device_env__ = True

    @dataclasses.dataclass(frozen=True)
    class View(FSMEnvView):
        avg_price: float

    def __init__(self, num_steps, network, seller_stage_handler=None, **batch_kwargs):
        buyers = [aid for aid, a in network.agents.items() if isinstance(a, BuyerAgent)]
        sellers = [aid for aid, a in network.agents.items() if isinstance(a, SellerAgent)]
        # `seller_stage_handler` (not in the example): an env handler for the Sellers stage, e.g.
        # a StageRule on the env-level avg_price words -- sellers re-price until it fires
        stages = [
            ph.FSMStage(stage_id="Buyers", next_stages=["Sellers"], acting_agents=buyers,
                        rewarded_agents=buyers),
            ph.FSMStage(stage_id="Sellers", acting_agents=sellers, rewarded_agents=sellers,
                        next_stages=["Buyers"] if seller_stage_handler is None else ["Buyers", "Sellers"],
                        handler=seller_stage_handler),
        ]
        super().__init__(num_steps, network, stages=stages, initial_stage="Sellers", **batch_kwargs)

    @property
    def avg_price(self):
        if not self.is_live:
            return 0.0
        lo = self.field(L.FIELD_ENV_STATE, np.int32, index=0).astype(np.uint32).astype(np.uint64)
        hi = self.field(L.FIELD_ENV_STATE, np.int32, index=1).astype(np.uint32).astype(np.uint64)
        v = ((hi << np.uint64(32)) | lo).view(np.float64)
        return v.item() if v.size == 1 else v

    def view(self, neighbour_id=None) -> "SimpleMarketEnv.View":
        return self.View(avg_price=self.avg_price, **super().view({}).__dict__)


def example_env(buyers=((0.2, 0.2, 0.2), (0.9, 1.0, 1.0), (0.9, 0.5, 0.5)), n_sellers: int = 2,
                num_steps: int = 10, enable_tracking: bool = False, **batch_kwargs) -> SimpleMarketEnv:
    """The cast of example_simple_market.py:9-30; buyers = (demand_prob, value low, value high)."""
    from phantom_b200.utils.samplers import UniformFloatSampler

    buyer_ids = [f"b{i + 1}" for i in range(len(buyers))]
    seller_ids = [f"s{i + 1}" for i in range(n_sellers)]
    agents = [BuyerAgent(b, p, supertype=BuyerSupertype(UniformFloatSampler(lo, hi)))
              for b, (p, lo, hi) in zip(buyer_ids, buyers)]
    agents += [SellerAgent(s) for s in seller_ids]
    network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking))
    network.add_connections_between(buyer_ids, seller_ids)
    return SimpleMarketEnv(num_steps=num_steps, network=network, **batch_kwargs)
