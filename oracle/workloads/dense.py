"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

BASELINE config C5: a dense-graph PhantomEnv that stresses the message queue and the
BatchResolver, written against the reference plugin API.  N strategic agents
(N = 128 in the benchmark), connected through `Network.add_connections_with_adjmat`
(network.py:145-177; complete graph by default), `BatchResolver(round_limit=2)`.

  acting   every agent with an action sends Signal(value) to EVERY neighbour, the value
           tailored to the receiver: value = rint(1000 a) + (7 s + 3 r) % 5  (s, r = sender /
           receiver slot), so every one of the N (N-1) messages is distinct.
  round 0  `handle_batch` is OVERRIDDEN (like the auction of
           examples/environments/digital_ads_market/digital_ads_market.py:429-510): one pass
           over the batch computes total, best (max) and the FIRST sender attaining it (batch
           order = global push order), then replies Ack(best) to that sender.
  round 1  the same override receives the Acks: counts them and sums their values.

Device twin: phantom_b200/csrc/fam_dense.cu (one thread block per env, mailbox in shared
memory).
"""
from __future__ import annotations

import numpy as np

N_AGENTS = 128
TYPE_SIGNAL, TYPE_ACK = 0, 1
MESSAGE_TYPE_IDS = {"Signal": 0, "Ack": 1}


def tailored(value: int, s: int, r: int) -> int:
    return value + (7 * s + 3 * r) % 5


def build(ph, *, n_agents: int = N_AGENTS, adjacency=None, num_steps: int = 8,
          round_limit=2, enable_tracking: bool = False):
    from ..phantom_oracle.spaces import Box

    @ph.msg_payload("DenseAgent", "DenseAgent")
    class Signal:
        value: int

    @ph.msg_payload("DenseAgent", "DenseAgent")
    class Ack:
        value: int

    ids = [f"N{i}" for i in range(n_agents)]
    slot = {a: i for i, a in enumerate(ids)}

    class DenseAgent(ph.StrategicAgent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.observation_space = Box(0.0, 1.0, (3,))
            self.action_space = Box(0.0, 1.0, (1,))
            self.reset()

        def reset(self):
            self.signal = 0
            self.total = 0
            self.best = 0
            self.best_sender = -1
            self.acks = 0
            self.ack_total = 0

        def pre_message_resolution(self, ctx):
            self.total = 0
            self.best = 0
            self.best_sender = -1
            self.acks = 0
            self.ack_total = 0

        def decode_action(self, ctx, action):
            self.signal = max(0, min(1000, int(round(action[0] * np.float32(1000.0)))))
            me = slot[self.id]
            out = [(n, Signal(tailored(self.signal, me, slot[n])))
                   for n in sorted(ctx.neighbour_ids, key=slot.get)]
            return out

        def handle_batch(self, ctx, batch):
            signals = [m for m in batch if isinstance(m.payload, Signal)]
            acks = [m for m in batch if isinstance(m.payload, Ack)]
            out = []
            if signals:
                self.total = sum(m.payload.value for m in signals)
                self.best = max(m.payload.value for m in signals)
                first = next(m for m in signals if m.payload.value == self.best)
                self.best_sender = slot[first.sender_id]
                out.append((first.sender_id, Ack(self.best)))
            if acks:
                self.acks += len(acks)
                self.ack_total += sum(m.payload.value for m in acks)
            return out

        def encode_observation(self, ctx):
            return np.array([self.total / 131072, self.best / 1024, self.acks / 128],
                            dtype=np.float32)

        def compute_reward(self, ctx):
            return self.ack_total / 1024

    agents = [DenseAgent(a) for a in ids]
    network = ph.Network(agents, ph.resolvers.BatchResolver(
        enable_tracking=enable_tracking, round_limit=round_limit))
    if adjacency is None:
        adjacency = np.ones((n_agents, n_agents), np.int64) - np.eye(n_agents, dtype=np.int64)
    network.add_connections_with_adjmat(ids, np.asarray(adjacency))
    env = ph.PhantomEnv(num_steps=num_steps, network=network)
    env.ids = ids
    return env


def random_adjacency(n: int, density: float, seed: int) -> np.ndarray:
    r = np.random.RandomState(seed)
    upper = np.triu((r.uniform(size=(n, n)) < density).astype(np.int64), 1)
    return upper + upper.T


def state(env):
    return np.array([[a.signal, a.total, a.best, a.best_sender, a.acks, a.ack_total]
                     for a in env.agents.values()], np.int64)
