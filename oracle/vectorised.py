"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

numpy restatements of the benchmark workloads, vectorised over envs, for parity checks at
sizes the object-level oracle (phantom_oracle) cannot reach in seconds (BASELINE's 65 536
envs).  Each class replays, per env, exactly the event order that the reference's routing
rules produce for that workload; tests/test_oracle_golden.py pins every one of them
against the object-level oracle and the reference-generated fixtures.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

from . import rng


class SupplyChainVec:
    """Supply chain (supply_chain.py:36-175) for E envs at once.

    Event order of one step, derived from phantom/env.py:239-303 and
    phantom/resolvers.py:128-163 on the star graph [SHOP, WAREHOUSE, CUST1..N]:
      sends   SHOP->WAREHOUSE StockRequest(ask), then CUSTi->SHOP OrderRequest(o_i)
      pre     shop.sales = shop.missed_sales = 0                  (supply_chain.py:93-96)
      round 0 receivers in first-arrival order: WAREHOUSE (echo), SHOP (fills o_1..o_N
              serially from the stock it had BEFORE this step's delivery)
      round 1 SHOP gets StockResponse(ask): stock = min(stock + ask, max_stock);
              customers get OrderResponse (no-op)
    """

    STREAM = 0

    def __init__(self, num_envs: int, seed: int, *, n_customers: int = 5, max_order: int = 5,
                 max_stock: int = 100, num_steps: int = 100, env_offset: int = 0):
        self.E, self.seed = num_envs, seed
        self.nc, self.max_order, self.max_stock, self.num_steps = (
            n_customers, max_order, max_stock, num_steps)
        self.env_ids = np.arange(env_offset, env_offset + num_envs, dtype=np.int64)
        z = lambda: np.zeros(num_envs, np.int64)
        self.stock, self.sales, self.missed, self.delivered = z(), z(), z(), z()
        self.step_no = z()
        self.episode = np.full(num_envs, -1, np.int64)

    def _obs(self) -> np.ndarray:
        cap = self.nc * self.max_order
        return np.stack(
            [self.stock / self.max_stock, self.sales / cap, self.missed / cap], axis=1
        ).astype(np.float32)

    def reset(self, mask: Optional[np.ndarray] = None) -> np.ndarray:
        m = np.ones(self.E, bool) if mask is None else mask.astype(bool)
        self.stock[m] = 0            # ShopAgent.reset clears only the stock
        self.step_no[m] = 0
        self.episode[m] += 1
        return self._obs()

    def step(self, actions: np.ndarray, action_mask: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
        a = np.asarray(actions, np.float32).reshape(self.E)
        has = np.ones(self.E, bool) if action_mask is None else action_mask.astype(bool).reshape(self.E)
        self.step_no += 1
        # decode_action: min(int(round(a)), max_stock - stock); round == rint on float32
        ask = np.minimum(np.rint(a).astype(np.int64), self.max_stock - self.stock)
        orders = rng.packed_randint_np(self.seed, self.env_ids, self.episode, self.step_no,
                                       self.STREAM, self.max_order, self.nc).astype(np.int64)
        self.sales[:] = 0
        self.missed[:] = 0
        for i in range(self.nc):  # serial order fill, round 0
            o = orders[:, i]
            short = o > self.stock
            self.missed += np.where(short, o - self.stock, 0)
            sold = np.where(short, self.stock, o)
            self.stock = np.where(short, 0, self.stock - o)
            self.sales += sold
        # round 1: delivery (only when the shop acted this step)
        self.delivered = np.where(has, ask, self.delivered)
        self.stock = np.where(has, np.minimum(self.stock + ask, self.max_stock), self.stock)
        reward = self.sales - 0.1 * self.stock
        at_max = self.step_no == self.num_steps
        zeros = np.zeros(self.E, np.uint8)
        return {
            "obs": self._obs(), "reward": reward, "term": zeros, "trunc": zeros.copy(),
            "all_term": zeros.copy(), "all_trunc": at_max.astype(np.uint8),
            "state": np.stack([self.stock, self.sales, self.missed, self.delivered], axis=1),
            "orders": orders, "ask": ask,
        }
