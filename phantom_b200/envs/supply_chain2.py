"""Multi-shop supply chain with agent supertypes on the device (family
PHX_FAMILY_SUPPLY_CHAIN2, csrc/fam_supply_chain2.cu) -- the reference tutorial's env
(docs/user/tutorial2.rst), 8 agents per env: WAREHOUSE, SHOP1, SHOP2, CUST1..5.  Same definition
as oracle/workloads/supply_chain2.py (which runs on the reference)."""
from __future__ import annotations

import dataclasses

import numpy as np

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box
from phantom_b200.utils.samplers import KIND_UNIFORM_FLOAT, Sampler, UniformFloatSampler

N_SHOPS, N_CUSTOMERS = 2, 5
MAX_ORDER, MAX_STOCK = 5, 100
MAX_EXCESS_STOCK_WEIGHT = 0.2
KIND_SHOP, KIND_FACTORY, KIND_CUSTOMER = 0, 1, 2


@ph.msg_payload("CustomerAgent", "ShopAgent")
class OrderRequest:
    size: int


@ph.msg_payload("ShopAgent", "CustomerAgent")
class OrderResponse:
    size: int


@ph.msg_payload("ShopAgent", "FactoryAgent")
class StockRequest:
    size: int


@ph.msg_payload("FactoryAgent", "ShopAgent")
class StockResponse:
    size: int


class FactoryAgent(ph.Agent):
    __phx_family__ = "supply_chain2"
    __phx_kind__ = KIND_FACTORY
    __phx_device_class__ = True


class CustomerAgent(ph.Agent):
    """Each step: OrderRequest(randint(max_order)) to a shop picked uniformly from shop_ids."""

    __phx_family__ = "supply_chain2"
    __phx_kind__ = KIND_CUSTOMER
    __phx_device_class__ = True

    def __init__(self, agent_id, shop_ids):
        super().__init__(agent_id)
        self.shop_ids = list(shop_ids)


class ShopAgent(ph.StrategicAgent):
    """reward = sales - type.excess_stock_weight * stock; the weight is an agent TYPE parameter,
    re-sampled per env and episode when its supertype field is a Sampler."""

    __phx_family__ = "supply_chain2"
    __phx_kind__ = KIND_SHOP
    __phx_device_class__ = True

    @dataclasses.dataclass
    class Supertype(ph.Supertype):
        excess_stock_weight: float = 0.1

    stock = device_column(0)
    sales = device_column(1)
    missed_sales = device_column(2)
    delivered_stock = device_column(3)

    def __init__(self, agent_id, factory_id):
        super().__init__(agent_id)
        self.factory_id = factory_id
        self.observation_space = Box(0.0, 1.0, (4,))
        self.action_space = Box(0.0, MAX_STOCK, (1,))

    @property
    def type(self):
        """agent.type as sampled for the current episode (float64 read back from the device)."""
        env = self._phx_env
        if env is None or not env.is_live:
            return (self.supertype or self.Supertype()).sample()
        lo = env.agent_column(self, 4).astype(np.uint32).astype(np.uint64)
        hi = env.agent_column(self, 5).astype(np.uint32).astype(np.uint64)
        w = ((hi << np.uint64(32)) | lo).view(np.float64)
        return self.Supertype(excess_stock_weight=w.item() if w.size == 1 else w)


def _collect(env, agents, spec) -> None:
    shops = [a for a in agents if isinstance(a, ShopAgent)]
    factories = [a for a in agents if isinstance(a, FactoryAgent)]
    customers = [a for a in agents if isinstance(a, CustomerAgent)]
    if len(factories) != 1 or not shops or not customers:
        raise NotLowerableError("supply-chain-2 device program: one factory, >= 1 shop / customer")
    slot = {a.id: a._phx_slot for a in agents}
    shop_ids = customers[0].shop_ids
    if any(c.shop_ids != shop_ids for c in customers) or len(shop_ids) > 6:
        raise NotLowerableError("customers must share one shop_ids list of at most 6 shops")
    spec.iparams[0], spec.iparams[1], spec.iparams[2] = MAX_ORDER, MAX_STOCK, len(shop_ids)
    for k, sid in enumerate(shop_ids):
        spec.iparams[3 + k] = slot[sid]
    spec.iparams[9] = len(customers)
    spec.fparams[0] = MAX_EXCESS_STOCK_WEIGHT
    for k, c in enumerate(customers):
        spec.agent_iparam[c._phx_slot][0] = k
    samplers = getattr(env, "_samplers", [])
    for s in shops:
        if s.factory_id != factories[0].id:
            raise NotLowerableError(f"shop '{s.id}' addresses unknown factory")
        i = s._phx_slot
        spec.agent_iparam[i][0] = factories[0]._phx_slot
        st = s.supertype if s.supertype is not None else s.Supertype()
        w = st.excess_stock_weight
        if isinstance(w, Sampler):
            kind, low, high = w.device_desc()
            if kind != KIND_UNIFORM_FLOAT:
                raise NotLowerableError("excess_stock_weight: only UniformFloatSampler is lowered")
            spec.agent_iparam[i][2] = next(k for k, x in enumerate(samplers) if x is w)
            spec.agent_fparam[i][0], spec.agent_fparam[i][1] = low, high
        else:
            spec.agent_iparam[i][2] = -1
            spec.agent_fparam[i][0] = float(w)


FAMILY = register(FamilyInfo(
    name="supply_chain2",
    family_id=L.FAMILY_SUPPLY_CHAIN2,
    payload_types=(OrderRequest, OrderResponse, StockRequest, StockResponse),
    obs_dim=4,
    act_dim=1,
    env_kinds=(L.ENV_BASE,),
    collect=_collect,
    trace_capacity=lambda env, agents: 4 * len(agents),
    supports_supertypes=True,
))


class SupplyChain2Env(ph.PhantomEnv):
    def __init__(self, n_shops: int = N_SHOPS, n_customers: int = N_CUSTOMERS, *,
                 num_steps: int = 100, agent_supertypes="default", enable_tracking: bool = False,
                 rates=None, shuffle_batches: bool = False, **batch_kwargs):
        """rates=(warehouse_rate, customer_rate): StochasticNetwork with those connection
        probabilities and ignore_connection_errors=True (oracle/workloads/supply_chain2.py)."""
        shop_ids = [f"SHOP{i + 1}" for i in range(n_shops)]
        customer_ids = [f"CUST{i + 1}" for i in range(n_customers)]
        agents = [FactoryAgent("WAREHOUSE")] + [ShopAgent(s, "WAREHOUSE") for s in shop_ids]
        agents += [CustomerAgent(c, shop_ids) for c in customer_ids]
        resolver = ph.resolvers.BatchResolver(enable_tracking=enable_tracking,
                                              shuffle_batches=shuffle_batches)
        if rates is None:
            network = ph.Network(agents, resolver)
            network.add_connections_between(shop_ids, ["WAREHOUSE"])
            network.add_connections_between(shop_ids, customer_ids)
        else:
            network = ph.StochasticNetwork(agents, resolver, ignore_connection_errors=True)
            network.add_connections_between(shop_ids, ["WAREHOUSE"], rate=rates[0])
            network.add_connections_between(shop_ids, customer_ids, rate=rates[1])
        if agent_supertypes == "default":  # the tutorial's training setup (tutorial2.rst:334-340)
            agent_supertypes = {
                s: {"excess_stock_weight": UniformFloatSampler(0.0, MAX_EXCESS_STOCK_WEIGHT)}
                for s in shop_ids}
        self.shop_ids, self.customer_ids = shop_ids, customer_ids
        super().__init__(num_steps=num_steps, network=network, agent_supertypes=agent_supertypes,
                         **batch_kwargs)
