"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The multi-shop supply chain with agent SUPERTYPES, modelled on the reference's tutorial
(docs/user/tutorial2.rst:90-128 two shops / customers pick a shop at random; :236-345 the
`excess_stock_weight` supertype sampled per episode with a UniformFloatSampler), written
against the reference plugin API.  8 agents per env: WAREHOUSE, SHOP1, SHOP2, CUST1..5 -- the
literal "8 agents/env" reading of BASELINE config C2.

  CustomerAgent   each step: OrderRequest(randint(max_order)) to a shop picked with
                  np.random.choice(shop_ids)                       (tutorial2.rst:90-97)
  ShopAgent       Supertype(excess_stock_weight); reward = sales - type.excess_stock_weight *
                  stock; obs = [stock/100, sales/25, missed/25, weight/0.2]  (:283-313)
  env             agent_supertypes={shop: {"excess_stock_weight": UniformFloatSampler(0, 0.2)}}
                  -> the env samples every Sampler at reset (phantom/env.py:212-216) and
                  Agent.reset() copies the values into agent.type (phantom/agents.py:166-168)

RNG contract: stream 0 order size (idx = customer ordinal), stream 4 shop choice (idx = customer
ordinal), stream 3 samplers at reset (step 0, idx = position in env._samplers).
Device twin: phantom_b200/csrc/fam_supply_chain2.cu.

Options exercising the "next" rows of SURVEY 8(f):
  rates=(warehouse_rate, customer_rate)  builds a StochasticNetwork(ignore_connection_errors=
      True) whose shop-warehouse / shop-customer connections exist with those probabilities,
      re-drawn at every reset (phantom/network.py:340-453); stream 5 (step 0, idx = position in
      _base_connections) replaces np.random.random().
  shuffle_batches=True  BatchResolver(shuffle_batches=True) (phantom/resolvers.py:150-151);
      np.random.shuffle is replaced by the contract's Fisher-Yates (oracle/harness.py).
"""
from __future__ import annotations

import dataclasses

import numpy as np

N_SHOPS, N_CUSTOMERS = 2, 5
MAX_ORDER, MAX_STOCK = 5, 100
MAX_EXCESS_STOCK_WEIGHT = 0.2
STREAM_ORDER, STREAM_SAMPLER, STREAM_SHOP_CHOICE = 0, 3, 4
MESSAGE_TYPE_IDS = {"OrderRequest": 0, "OrderResponse": 1, "StockRequest": 2, "StockResponse": 3}


STREAM_CONNECTIVITY = 5


def build(ph, streams, sampler_factory, *, n_shops: int = N_SHOPS, n_customers: int = N_CUSTOMERS,
          num_steps: int = 100, enable_tracking: bool = False, rates=None,
          shuffle_batches: bool = False):
    """streams: {STREAM_ORDER: StepStream, STREAM_SHOP_CHOICE: StepStream};
    sampler_factory(low, high) -> a Sampler of the API in use whose sample() follows the
    contract (oracle: ContractUniformFloatSampler; reference: its own UniformFloatSampler with
    np.random.uniform patched)."""
    from ..phantom_oracle.spaces import Box

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

    shop_ids = [f"SHOP{i + 1}" for i in range(n_shops)]
    customer_ids = [f"CUST{i + 1}" for i in range(n_customers)]

    class FactoryAgent(ph.Agent):
        @ph.agents.msg_handler(StockRequest)
        def on_stock_request(self, ctx, message):
            return [(message.sender_id, StockResponse(message.payload.size))]

    class CustomerAgent(ph.Agent):
        def __init__(self, agent_id, shop_ids):
            super().__init__(agent_id)
            self.shop_ids = shop_ids

        @ph.agents.msg_handler(OrderResponse)
        def on_order_response(self, ctx, message):
            return None

        def generate_messages(self, ctx):
            size = streams[STREAM_ORDER].randint(MAX_ORDER)
            shop = self.shop_ids[streams[STREAM_SHOP_CHOICE].randint(len(self.shop_ids))]
            return [(shop, OrderRequest(size))]

    class ShopAgent(ph.StrategicAgent):
        @dataclasses.dataclass
        class Supertype(ph.Supertype):
            excess_stock_weight: float = 0.1

        def __init__(self, agent_id, factory_id):
            super().__init__(agent_id)
            self.factory_id = factory_id
            self.stock = self.sales = self.missed_sales = 0
            self.observation_space = Box(0.0, 1.0, (4,))
            self.action_space = Box(0.0, MAX_STOCK, (1,))

        def pre_message_resolution(self, ctx):
            self.sales = 0
            self.missed_sales = 0

        @ph.agents.msg_handler(StockResponse)
        def on_stock_response(self, ctx, message):
            self.delivered_stock = message.payload.size
            self.stock = min(self.stock + self.delivered_stock, MAX_STOCK)

        @ph.agents.msg_handler(OrderRequest)
        def on_order_request(self, ctx, message):
            wanted = message.payload.size
            if wanted > self.stock:
                self.missed_sales += wanted - self.stock
                sold, self.stock = self.stock, 0
            else:
                sold = wanted
                self.stock -= wanted
            self.sales += sold
            return [(message.sender_id, OrderResponse(sold))]

        def encode_observation(self, ctx):
            cap = n_customers * MAX_ORDER
            return np.array(
                [self.stock / MAX_STOCK, self.sales / cap, self.missed_sales / cap,
                 self.type.excess_stock_weight / MAX_EXCESS_STOCK_WEIGHT], dtype=np.float32)

        def decode_action(self, ctx, action):
            ask = min(int(round(action[0])), MAX_STOCK - self.stock)
            return [(self.factory_id, StockRequest(ask))]

        def compute_reward(self, ctx):
            return self.sales - self.type.excess_stock_weight * self.stock

        def reset(self):
            super().reset()  # self.type = self.supertype.sample()
            self.stock = 0

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
    supertypes = {s: ShopAgent.Supertype(sampler_factory(0.0, MAX_EXCESS_STOCK_WEIGHT))
                  for s in shop_ids}
    env = ph.PhantomEnv(num_steps=num_steps, network=network, agent_supertypes=supertypes)
    env.shop_ids, env.customer_ids = shop_ids, customer_ids
    return env


def state(env):
    rows = []
    for s in env.shop_ids:
        a = env.agents[s]
        rows.append([a.stock, a.sales, a.missed_sales, getattr(a, "delivered_stock", 0)])
    return np.array(rows, np.int64)


def adjacency(env):
    """Current graph as uint8 [n, n] in agent order (nx.DiGraph or the oracle's _DiGraph)."""
    ids = list(env.agent_ids)
    g = env.network.graph
    return np.array([[1 if g.has_edge(u, v) else 0 for v in ids] for u in ids], np.uint8)


def weights(env):
    return np.array([env.agents[s].type.excess_stock_weight for s in env.shop_ids], np.float64)
