"""A user-defined env class on a user-written device program (test fixture for SURVEY 8b: the
plugin API is "bring your own agent classes").  `build_device` is what a phantom_b200 user
writes -- agent classes that name the family of auction_game.cu; `build_reference` is the same
env with Python handlers against the reference plugin API (run on the oracle port / reference)."""
import os

import numpy as np

WINS_TO_RETIRE = 5
N_BIDDERS = 3
HERE = os.path.dirname(os.path.abspath(__file__))


def build_device(n_bidders=N_BIDDERS, **batch_kwargs):
    import phantom_b200 as ph
    from phantom_b200.agents import device_column
    from phantom_b200.families import REGISTRY, register_user_family
    from phantom_b200.spaces import Box

    if "auction_game" not in REGISTRY:
        @ph.msg_payload("Bidder", "Book")
        class Bid:
            price: int

        @ph.msg_payload("Book", "Bidder")
        class Ack:
            rank: int
            best: int

        class Bidder(ph.StrategicAgent):
            __phx_family__ = "auction_game"
            __phx_kind__ = 0
            __phx_device_class__ = True
            last_rank = device_column(0)
            wins = device_column(1)
            last_best = device_column(2)

            def __init__(self, agent_id, book_id):
                super().__init__(agent_id)
                self.book_id = book_id
                self.observation_space = Box(0.0, 1.0, (3,))
                self.action_space = Box(0.0, 1.0, (1,))

        class Book(ph.Agent):
            __phx_family__ = "auction_game"
            __phx_kind__ = 1
            __phx_device_class__ = True
            count = device_column(0)
            best = device_column(1)
            total = device_column(2)

        def collect(env, agents, spec):
            spec.iparams[0] = WINS_TO_RETIRE
            for a in agents:
                if isinstance(a, Bidder):
                    spec.agent_iparam[a._phx_slot][0] = env.agents[a.book_id]._phx_slot

        info = register_user_family("auction_game", os.path.join(HERE, "auction_game.cu"),
                                    (Bid, Ack), obs_dim=3, act_dim=1, collect=collect)
        info.classes = (Bidder, Book)
    Bidder, Book = REGISTRY["auction_game"].classes
    ids = [f"B{i + 1}" for i in range(n_bidders)]
    agents = [Bidder(b, "BOOK") for b in ids] + [Book("BOOK")]
    net = ph.Network(agents, ph.resolvers.BatchResolver(
        enable_tracking=batch_kwargs.pop("enable_tracking", False)))
    net.add_connections_between(["BOOK"], ids)
    return ph.PhantomEnv(num_steps=batch_kwargs.pop("num_steps", 40), network=net, **batch_kwargs)


def build_reference(ph, num_steps=40, n_bidders=N_BIDDERS):
    """The same env with Python handlers (reference plugin API)."""
    from oracle.phantom_oracle.spaces import Box

    @ph.msg_payload("Bidder", "Book")
    class Bid:
        price: int

    @ph.msg_payload("Book", "Bidder")
    class Ack:
        rank: int
        best: int

    class Bidder(ph.StrategicAgent):
        def __init__(self, agent_id, book_id):
            super().__init__(agent_id)
            self.book_id = book_id
            self.observation_space = Box(0.0, 1.0, (3,))
            self.action_space = Box(0.0, 1.0, (1,))
            self.reset()

        def reset(self):
            self.last_rank = self.wins = self.last_best = 0

        def decode_action(self, ctx, action):
            return [(self.book_id, Bid(int(round(action[0] * np.float32(100.0)))))]

        @ph.agents.msg_handler(Ack)
        def on_ack(self, ctx, message):
            self.last_rank = message.payload.rank
            self.last_best = message.payload.best
            if message.payload.rank == 1:
                self.wins += 1

        def encode_observation(self, ctx):
            return np.array([self.last_rank / 4, self.wins / 100, self.last_best / 100], np.float32)

        def compute_reward(self, ctx):
            return (4 - self.last_rank) / 4

        def is_terminated(self, ctx):
            return self.wins >= WINS_TO_RETIRE

    class Book(ph.Agent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.reset()

        def reset(self):
            self.count = self.best = self.total = 0

        def pre_message_resolution(self, ctx):
            self.count = self.best = self.total = 0

        @ph.agents.msg_handler(Bid)
        def on_bid(self, ctx, message):
            self.count += 1
            self.total += message.payload.price
            self.best = max(self.best, message.payload.price)
            return [(message.sender_id, Ack(self.count, self.best))]

    ids = [f"B{i + 1}" for i in range(n_bidders)]
    agents = [Bidder(b, "BOOK") for b in ids] + [Book("BOOK")]
    net = ph.Network(agents, ph.resolvers.BatchResolver())
    net.add_connections_between(["BOOK"], ids)
    return ph.PhantomEnv(num_steps=num_steps, network=net)
