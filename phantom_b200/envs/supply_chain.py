"""Supply-chain env on the device (family PHX_FAMILY_SUPPLY_CHAIN, csrc/fam_supply_chain.cu).

Drop-in for the reference example
/root/reference/examples/environments/supply_chain/supply_chain.py: same payload classes,
agent classes, constructor arguments, agent ids and network wiring; the handler bodies live
in the fused kernel.  `SupplyChainEnv(num_envs=65536)` is BASELINE config C2.
"""
from __future__ import annotations

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box

NUM_EPISODE_STEPS = 100
NUM_CUSTOMERS = 5
CUSTOMER_MAX_ORDER_SIZE = 5
SHOP_MAX_STOCK = 100

KIND_SHOP, KIND_FACTORY, KIND_CUSTOMER = 0, 1, 2
FIELD_SHOP_STATE = L.FIELD_FAMILY + 0  # int32 [E,4]: stock, sales, missed_sales, delivered_stock


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
    """Echoes every StockRequest back as a StockResponse (unlimited production)."""

    __phx_family__ = "supply_chain"
    __phx_kind__ = KIND_FACTORY
    __phx_device_class__ = True

    def __init__(self, agent_id: str):
        super().__init__(agent_id)


class CustomerAgent(ph.Agent):
    """Sends one OrderRequest(randint(max_order)) to its shop every step."""

    __phx_family__ = "supply_chain"
    __phx_kind__ = KIND_CUSTOMER
    __phx_device_class__ = True

    def __init__(self, agent_id: ph.AgentID, shop_id: ph.AgentID):
        super().__init__(agent_id)
        self.shop_id = shop_id


class ShopAgent(ph.StrategicAgent):
    """Holds stock, fills customer orders serially, restocks from the factory.
    obs = [stock/max_stock, sales/cap, missed/cap]; action = [restock quantity];
    reward = sales - 0.1 * stock."""

    __phx_family__ = "supply_chain"
    __phx_kind__ = KIND_SHOP
    __phx_device_class__ = True

    stock = device_column(0)
    sales = device_column(1)
    missed_sales = device_column(2)
    delivered_stock = device_column(3)

    def __init__(self, agent_id: str, factory_id: str):
        super().__init__(agent_id)
        self.factory_id = factory_id
        self.observation_space = Box(low=0.0, high=1.0, shape=(3,))
        self.action_space = Box(low=0.0, high=SHOP_MAX_STOCK, shape=(1,))


def _collect(env, agents, spec) -> None:
    shops = [a for a in agents if isinstance(a, ShopAgent)]
    factories = [a for a in agents if isinstance(a, FactoryAgent)]
    if len(shops) != 1 or len(factories) != 1:
        raise NotLowerableError("supply-chain device program: exactly one ShopAgent and one "
                                "FactoryAgent per env")
    shop, factory = shops[0], factories[0]
    if shop.factory_id != factory.id:
        raise NotLowerableError(f"ShopAgent.factory_id '{shop.factory_id}' is not the factory")
    for a in agents:
        if isinstance(a, CustomerAgent) and a.shop_id != shop.id:
            raise NotLowerableError(f"CustomerAgent '{a.id}' addresses unknown shop '{a.shop_id}'")
    spec.iparams[0] = int(getattr(env, "max_order", CUSTOMER_MAX_ORDER_SIZE))
    spec.iparams[1] = int(getattr(env, "max_stock", SHOP_MAX_STOCK))
    ordinal = 0
    for i, a in enumerate(agents):
        if isinstance(a, CustomerAgent):
            spec.agent_iparam[i][0] = shop._phx_slot
            spec.agent_iparam[i][1] = ordinal  # RNG idx: k-th customer in agent order
            ordinal += 1
        elif isinstance(a, ShopAgent):
            spec.agent_iparam[i][0] = factory._phx_slot


def _fast_column(env, agent, word, value=None):
    """Shop state of the thread-per-env kernel: one int4 per env."""
    import numpy as np

    col = env.field(FIELD_SHOP_STATE, np.int32, width=4)
    if value is None:
        return col[:, word]
    col[:, word] = value
    env.set_field(FIELD_SHOP_STATE, col)


FAMILY = register(FamilyInfo(
    name="supply_chain",
    family_id=L.FAMILY_SUPPLY_CHAIN,
    payload_types=(OrderRequest, OrderResponse, StockRequest, StockResponse),
    obs_dim=3,
    act_dim=1,
    env_kinds=(L.ENV_BASE,),
    collect=_collect,
    trace_capacity=lambda env, agents: 2 * len(agents),
    fast_column=_fast_column,
))


class SupplyChainEnv(ph.PhantomEnv):
    """1 shop (strategic) + 1 factory + N customers, star on the shop, 100-step episodes.

    `SupplyChainEnv()` is the reference env; keyword arguments select the batch size,
    device, RNG base seed and kernel variant."""

    def __init__(self, n_customers: int = NUM_CUSTOMERS, *, num_steps: int = NUM_EPISODE_STEPS,
                 enable_tracking: bool = False, **batch_kwargs):
        factory_id = "WAREHOUSE"
        customer_ids = [f"CUST{i + 1}" for i in range(n_customers)]
        shop_id = "SHOP"

        factory_agent = FactoryAgent(factory_id)
        customer_agents = [CustomerAgent(cid, shop_id=shop_id) for cid in customer_ids]
        shop_agent = ShopAgent(shop_id, factory_id=factory_id)

        agents = [shop_agent, factory_agent] + customer_agents
        network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking))
        network.add_connection(shop_id, factory_id)
        network.add_connections_between([shop_id], customer_ids)

        self.max_order = CUSTOMER_MAX_ORDER_SIZE
        self.max_stock = SHOP_MAX_STOCK
        super().__init__(num_steps=num_steps, network=network, **batch_kwargs)
