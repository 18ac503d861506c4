"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

BASELINE config C4: a leader-follower pricing game under StackelbergEnv, written against the
reference plugin API.  Agent order [LEADER, F1..Fn], star on the leader.

  odd steps   the leader acts: Price(ticks) to every follower, who remember it.
  even steps  the followers act: Demand(qty) to the leader, who serves them first come first
              served from CAPACITY units (order dependent) and answers Ack(filled) in the next
              resolver round; a follower books utility filled * (value - price).

Prices are integer ticks.  Device twin: phantom_b200/csrc/fam_stackelberg.cu.
"""
from __future__ import annotations

import numpy as np

N_FOLLOWERS = 3
CAPACITY = 15
STREAM_FOLLOWER_VALUE = 2
KIND_LEADER, KIND_FOLLOWER = 0, 1
TYPE_PRICE, TYPE_DEMAND, TYPE_ACK = 0, 1, 2
MESSAGE_TYPE_IDS = {"Price": 0, "Demand": 1, "Ack": 2}


def build(ph, stream, *, n_followers: int = N_FOLLOWERS, num_steps: int = 100,
          enable_tracking: bool = False):
    """`stream`: oracle.rng.StepStream for STREAM_FOLLOWER_VALUE (valuations at reset)."""
    from ..phantom_oracle.spaces import Box

    @ph.msg_payload("LeaderAgent", "FollowerAgent")
    class Price:
        ticks: int

    @ph.msg_payload("FollowerAgent", "LeaderAgent")
    class Demand:
        qty: int

    @ph.msg_payload("LeaderAgent", "FollowerAgent")
    class Ack:
        filled: int

    follower_ids = [f"F{i + 1}" for i in range(n_followers)]

    class LeaderAgent(ph.StrategicAgent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.observation_space = Box(0.0, 1.0, (2,))
            self.action_space = Box(0.0, 1.0, (1,))
            self.reset()

        def reset(self):
            self.price = 0
            self.remaining = CAPACITY
            self.revenue_round = 0
            self.demand_round = 0

        def pre_message_resolution(self, ctx):
            if ctx.env_view.current_step % 2 == 0:  # followers' turn: a new selling round
                self.remaining = CAPACITY
                self.revenue_round = 0
                self.demand_round = 0

        def decode_action(self, ctx, action):
            self.price = max(0, min(100, int(round(action[0] * np.float32(100.0)))))
            return [(f, Price(self.price)) for f in follower_ids if f in ctx]

        @ph.agents.msg_handler(Demand)
        def on_demand(self, ctx, message):
            filled = min(message.payload.qty, self.remaining)
            self.remaining -= filled
            self.revenue_round += filled * self.price
            self.demand_round += message.payload.qty
            return [(message.sender_id, Ack(filled))]

        def encode_observation(self, ctx):
            return np.array([self.demand_round / 30, self.remaining / CAPACITY], dtype=np.float32)

        def compute_reward(self, ctx):
            return self.revenue_round / 100

    class FollowerAgent(ph.StrategicAgent):
        def __init__(self, agent_id, leader_id):
            super().__init__(agent_id)
            self.leader_id = leader_id
            self.observation_space = Box(0.0, 1.0, (2,))
            self.action_space = Box(0.0, 1.0, (1,))
            self.value = 0
            self.seen_price = 0
            self.last_filled = 0
            self.utility_round = 0

        def reset(self):
            self.value = 50 + stream.randint(51)
            self.seen_price = 0
            self.last_filled = 0
            self.utility_round = 0

        def pre_message_resolution(self, ctx):
            if ctx.env_view.current_step % 2 == 0:
                self.last_filled = 0
                self.utility_round = 0

        @ph.agents.msg_handler(Price)
        def on_price(self, ctx, message):
            self.seen_price = message.payload.ticks

        def decode_action(self, ctx, action):
            qty = max(0, min(10, int(round(action[0] * np.float32(10.0)))))
            return [(self.leader_id, Demand(qty))]

        @ph.agents.msg_handler(Ack)
        def on_ack(self, ctx, message):
            self.last_filled = message.payload.filled
            self.utility_round = message.payload.filled * (self.value - self.seen_price)

        def encode_observation(self, ctx):
            return np.array([self.seen_price / 100, self.last_filled / 10], dtype=np.float32)

        def compute_reward(self, ctx):
            return self.utility_round / 100

    agents = [LeaderAgent("LEADER")] + [FollowerAgent(f, "LEADER") for f in follower_ids]
    network = ph.Network(agents, ph.resolvers.BatchResolver(enable_tracking=enable_tracking))
    network.add_connections_between(["LEADER"], follower_ids)
    env = ph.StackelbergEnv(num_steps, network, ["LEADER"], follower_ids)
    env.follower_ids = follower_ids
    return env


def state(env):
    rows = []
    for a in env.agents.values():
        if type(a).__name__ == "LeaderAgent":
            rows.append([a.price, a.remaining, a.revenue_round, a.demand_round])
        else:
            rows.append([a.value, a.seen_price, a.last_filled, a.utility_round])
    return np.array(rows, np.int64)
