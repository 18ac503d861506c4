"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

Supply-chain workload (BASELINE configs C1/C2) restated against the reference plugin API.

Follows /root/reference/examples/environments/supply_chain/supply_chain.py:
  :16-33   the four payload types and their sender/receiver whitelists
  :36-45   FactoryAgent -- echoes a StockRequest back as a StockResponse
  :48-67   CustomerAgent -- one OrderRequest(randint(max_order)) per step, ignores replies
  :70-150  ShopAgent -- stock / sales / missed_sales bookkeeping, obs, reward, action
  :153-175 SupplyChainEnv -- agent order [SHOP, WAREHOUSE, CUST1..N], star on SHOP

The only change is the RNG call site (:64): instead of process-global
`np.random.randint`, the customer pulls from the env's counter-based stream
(oracle/rng.py packed draws, stream 0, draw i = customer index; `order_stream` below).  tests/test_oracle_golden.py proves this
file == the unmodified example with `np.random.randint` patched to the same stream.

Device twin: phantom_b200/csrc/fam_supply_chain.cuh.
"""
from __future__ import annotations

import numpy as np

STREAM_CUSTOMER_ORDER = 0

# slot / type numbering shared with the device program
SLOT_SHOP, SLOT_FACTORY, SLOT_FIRST_CUSTOMER = 0, 1, 2
TYPE_ORDER_REQUEST, TYPE_ORDER_RESPONSE, TYPE_STOCK_REQUEST, TYPE_STOCK_RESPONSE = 0, 1, 2, 3
PAYLOAD_TYPE_IDS = {
    "OrderRequest": TYPE_ORDER_REQUEST,
    "OrderResponse": TYPE_ORDER_RESPONSE,
    "StockRequest": TYPE_STOCK_REQUEST,
    "StockResponse": TYPE_STOCK_RESPONSE,
}


def order_stream(seed: int, env: int, *, n_customers: int = 5, max_order: int = 5):
    """The contract stream of the customers' order sizes for global env index `env`:
    K = n_customers packed draws of randint(max_order) per step (oracle/rng.py)."""
    from .. import rng

    return rng.PackedStream(seed, env, STREAM_CUSTOMER_ORDER, max_order, n_customers)


def build(ph, stream, *, n_customers: int = 5, max_order: int = 5, max_stock: int = 100,
          num_steps: int = 100, enable_tracking: bool = False):
    """Return a supply-chain PhantomEnv built with API module `ph`.

    `stream`: `order_stream(...)` (or anything with .randint(n))."""

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
        @ph.agents.msg_handler(StockRequest)
        def on_stock_request(self, ctx, message):
            return [(message.sender_id, StockResponse(message.payload.size))]

    class CustomerAgent(ph.Agent):
        def __init__(self, agent_id, shop_id):
            super().__init__(agent_id)
            self.shop_id = shop_id

        @ph.agents.msg_handler(OrderResponse)
        def on_order_response(self, ctx, message):
            return None

        def generate_messages(self, ctx):
            return [(self.shop_id, OrderRequest(stream.randint(max_order)))]

    class ShopAgent(ph.StrategicAgent):
        def __init__(self, agent_id, factory_id):
            super().__init__(agent_id)
            self.factory_id = factory_id
            self.stock = 0
            self.sales = 0
            self.missed_sales = 0

        def pre_message_resolution(self, ctx):
            self.sales = 0
            self.missed_sales = 0

        @ph.agents.msg_handler(StockResponse)
        def on_stock_response(self, ctx, message):
            self.delivered_stock = message.payload.size
            self.stock = min(self.stock + self.delivered_stock, max_stock)

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
            cap = n_customers * max_order
            return np.array(
                [self.stock / max_stock, self.sales / cap, self.missed_sales / cap],
                dtype=np.float32,
            )

        def decode_action(self, ctx, action):
            ask = min(int(round(action[0])), max_stock - self.stock)
            return [(self.factory_id, StockRequest(ask))]

        def compute_reward(self, ctx):
            return self.sales - 0.1 * self.stock

        def reset(self):
            self.stock = 0  # sales / missed_sales deliberately survive (supply_chain.py:149-150)

    customers = [f"CUST{i + 1}" for i in range(n_customers)]
    agents = [ShopAgent("SHOP", "WAREHOUSE"), FactoryAgent("WAREHOUSE")]
    agents += [CustomerAgent(c, "SHOP") for c in customers]
    network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking))
    network.add_connection("SHOP", "WAREHOUSE")
    network.add_connections_between(["SHOP"], customers)
    return ph.PhantomEnv(num_steps=num_steps, network=network)


def shop_state(env) -> tuple:
    s = env.agents["SHOP"]
    return (s.stock, s.sales, s.missed_sales, getattr(s, "delivered_stock", 0))
