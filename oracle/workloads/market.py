"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

BASELINE config C3: a three-stage FiniteStateMachineEnv market, written against the reference
plugin API (modelled on examples/environments/simple_market/ and the digital-ads auction;
SURVEY.md 8d).  32 agents: N_MAKERS makers (strategic), N_TAKERS takers (strategic), one
clearing agent (non-strategic).  Agent order: makers, takers, clearing.

  stage MAKER     acting: makers.   Each maker quotes Quote(price) to every taker.  A taker keeps
                  the best (lowest) quote it hears, first arrival winning ties, ignoring makers
                  whose start-of-step view shows no inventory (AgentView, views.py:20-24).
  stage TAKER     acting: takers.   A taker that decides to buy sends Order(maker, price) to the
                  clearing agent, which accepts orders first come first served up to
                  MAKER_CAPACITY units per maker and cycle (order dependent).
  stage CLEARING  acting: clearing agent (generate_messages).  Sends Fill(units, notional) to
                  every maker and Fill(accepted, 0) to every taker; they settle.
                  rewarded: makers + takers (rewards reach the takers with their NEXT
                  observation through the FSM reward cache, fsm.py:378).
  MAKER -> TAKER -> CLEARING -> MAKER, handler-less (deterministic).

Prices are integer ticks (0..100) so that everything but the final obs/reward scaling is exact
integer arithmetic.  Device twin: phantom_b200/csrc/fam_market.cu.
"""
from __future__ import annotations

import dataclasses

import numpy as np

N_MAKERS, N_TAKERS = 7, 24
MAKER_INVENTORY = 12   # units at reset; a maker terminates when it is sold out
MAKER_CAPACITY = 3     # units per maker and cycle
NO_QUOTE = 1 << 20
STREAM_TAKER_VALUE = 1

STAGES = ("MAKER", "TAKER", "CLEARING")
KIND_MAKER, KIND_TAKER, KIND_CLEARING = 0, 1, 2
TYPE_QUOTE, TYPE_ORDER, TYPE_FILL = 0, 1, 2


def build(ph, stream, *, n_makers: int = N_MAKERS, n_takers: int = N_TAKERS,
          num_steps: int = 99, enable_tracking: bool = False, shuffle_batches: bool = False):
    """`stream`: oracle.rng.StepStream for STREAM_TAKER_VALUE (taker valuations at reset)."""
    from ..phantom_oracle.spaces import Box, Discrete

    @ph.msg_payload("MakerAgent", "TakerAgent")
    class Quote:
        price: int

    @ph.msg_payload("TakerAgent", "ClearingAgent")
    class Order:
        maker: int  # maker ordinal
        price: int

    @ph.msg_payload("ClearingAgent", ["MakerAgent", "TakerAgent"])
    class Fill:
        units: int
        notional: int

    @dataclasses.dataclass(frozen=True)
    class MakerView(ph.AgentView):
        inventory: int

    maker_ids = [f"M{i + 1}" for i in range(n_makers)]
    taker_ids = [f"T{i + 1}" for i in range(n_takers)]
    maker_ordinal = {m: i for i, m in enumerate(maker_ids)}
    taker_ordinal = {t: i for i, t in enumerate(taker_ids)}

    class MakerAgent(ph.StrategicAgent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.observation_space = Box(0.0, 1.0, (2,))
            self.action_space = Box(0.0, 1.0, (1,))
            self.inventory = MAKER_INVENTORY
            self.cash = 0
            self.last_price = 0
            self.last_notional = 0

        def view(self, neighbour_id=None):
            return MakerView(self.inventory)

        def decode_action(self, ctx, action):
            self.last_price = max(0, min(100, int(round(action[0] * np.float32(100.0)))))
            return [(t, Quote(self.last_price)) for t in taker_ids if t in ctx]

        @ph.agents.msg_handler(Fill)
        def on_fill(self, ctx, message):
            self.inventory -= message.payload.units
            self.cash += message.payload.notional
            self.last_notional = message.payload.notional

        def encode_observation(self, ctx):
            return np.array([self.inventory / MAKER_INVENTORY, self.last_price / 100],
                            dtype=np.float32)

        def compute_reward(self, ctx):
            return self.last_notional / 100

        def is_terminated(self, ctx):
            return self.inventory <= 0

        def reset(self):
            self.inventory = MAKER_INVENTORY
            self.cash = 0
            self.last_price = 0
            self.last_notional = 0

    class TakerAgent(ph.StrategicAgent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.observation_space = Box(0.0, 1.0, (3,))
            self.action_space = Discrete(2)
            self.value = 0
            self.best_price = NO_QUOTE
            self.best_maker = -1
            self.holdings = 0
            self.last_surplus = 0

        def pre_message_resolution(self, ctx):
            if ctx.env_view.stage == "MAKER":  # a new cycle: forget last cycle's quotes
                self.best_price = NO_QUOTE
                self.best_maker = -1
                self.last_surplus = 0

        @ph.agents.msg_handler(Quote)
        def on_quote(self, ctx, message):
            if ctx[message.sender_id].inventory <= 0:  # start-of-step view of the maker
                return
            if message.payload.price < self.best_price:
                self.best_price = message.payload.price
                self.best_maker = maker_ordinal[message.sender_id]

        def decode_action(self, ctx, action):
            if int(action) == 1 and self.best_maker >= 0:
                return [("CLEARING", Order(self.best_maker, self.best_price))]
            return []

        @ph.agents.msg_handler(Fill)
        def on_fill(self, ctx, message):
            if message.payload.units > 0:
                self.holdings += 1
                self.last_surplus = self.value - self.best_price

        def encode_observation(self, ctx):
            quote = 100 if self.best_maker < 0 else self.best_price
            return np.array([self.value / 100, quote / 100, self.holdings / 33],
                            dtype=np.float32)

        def compute_reward(self, ctx):
            return self.last_surplus / 100

        def reset(self):
            self.value = stream.randint(101)
            self.best_price = NO_QUOTE
            self.best_maker = -1
            self.holdings = 0
            self.last_surplus = 0

    class ClearingAgent(ph.Agent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.units = [0] * n_makers
            self.notional = [0] * n_makers
            self.accepted = [0] * n_takers

        @ph.agents.msg_handler(Order)
        def on_order(self, ctx, message):
            m = message.payload.maker
            if self.units[m] < MAKER_CAPACITY:  # first come, first served
                self.units[m] += 1
                self.notional[m] += message.payload.price
                self.accepted[taker_ordinal[message.sender_id]] = 1

        def generate_messages(self, ctx):
            if ctx.env_view.stage != "CLEARING":
                return []
            out = [(m, Fill(self.units[i], self.notional[i]))
                   for i, m in enumerate(maker_ids) if m in ctx]
            out += [(t, Fill(self.accepted[i], 0)) for i, t in enumerate(taker_ids) if t in ctx]
            return out

        def post_message_resolution(self, ctx):
            if ctx.env_view.stage == "CLEARING":  # book settled
                self.units = [0] * n_makers
                self.notional = [0] * n_makers
                self.accepted = [0] * n_takers

        def reset(self):
            self.units = [0] * n_makers
            self.notional = [0] * n_makers
            self.accepted = [0] * n_takers

    agents = [MakerAgent(m) for m in maker_ids] + [TakerAgent(t) for t in taker_ids]
    agents.append(ClearingAgent("CLEARING"))
    network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking,
                                                            shuffle_batches=shuffle_batches))
    network.add_connections_between(maker_ids, taker_ids)
    network.add_connections_between(["CLEARING"], maker_ids + taker_ids)
    everyone = maker_ids + taker_ids
    env = ph.FiniteStateMachineEnv(
        num_steps=num_steps, network=network, initial_stage="MAKER",
        stages=[
            ph.FSMStage("MAKER", acting_agents=maker_ids, rewarded_agents=[], next_stages=["TAKER"]),
            ph.FSMStage("TAKER", acting_agents=taker_ids, rewarded_agents=[], next_stages=["CLEARING"]),
            ph.FSMStage("CLEARING", acting_agents=["CLEARING"], rewarded_agents=everyone,
                        next_stages=["MAKER"]),
        ])
    env.maker_ids, env.taker_ids = maker_ids, taker_ids
    return env
