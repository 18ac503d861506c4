"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The reference's second example environment, examples/environments/simple_market/ -- a
two-stage FiniteStateMachineEnv ("Sellers" -> "Buyers" -> "Sellers" ...) with an ENV-LEVEL
`post_message_resolution` override and a custom `EnvView` field:

  market_agents.py:33-85    BuyerAgent   keeps the last price heard from every seller; action 1
                            = buy one unit from a cheapest seller (ties: random.choice);
                            obs [min price, demand ~ binomial(1, demand_prob), type.value];
                            reward accumulated in decode_action, cleared by compute_reward
  market_agents.py:92-129   SellerAgent  action = price, sent to every neighbour; books revenue
                            and volume per Order; obs [volume, env_view.avg_price]
  simple_mkt_env.py:9-58    SimpleMarketEnv: stages, `view()` adds avg_price to the FSM env view,
                            `post_message_resolution()` sets avg_price = np.mean(seller prices)

Two ways to get an env:
  build_reference(...)  imports the UNMODIFIED example modules from /root/reference (build
                        container only) -- this is what generates tests/golden/simple_market_*.
  build(ph, ...)        the same definitions restated against the plugin API module `ph`
                        (the oracle port), for the GPU box and for other seeds / sizes.

RNG call sites and their contract replacement (oracle/rng.py, 24-bit draws; `contract_rng`):
  stream 6  reset draws, step 0, idx = k-th draw of the reset in agent order:
            UniformFloatSampler.sample (samplers.py:142: np.random.uniform(low, high)) for a
            buyer's type.value, Box.sample() for a seller's initial price (market_agents.py:121)
  stream 7  random.choice(min_sellers) (market_agents.py:58): index randint(len), idx = buyer ordinal
  stream 8  np.random.binomial(1, p) (market_agents.py:68): 1 iff d24 * 2^-24 < p, idx = buyer ordinal
Device twin: phantom_b200/csrc/fam_simple_market.cu.
"""
from __future__ import annotations

import contextlib
import dataclasses
import os
import random
import sys

import numpy as np

from .. import rng

STREAM_RESET, STREAM_CHOICE, STREAM_DEMAND = 6, 7, 8
OBS_DIM = 3

# the example script's cast (example_simple_market.py:9-16): (demand_prob, value low, value high)
EXAMPLE_BUYERS = ((0.2, 0.2, 0.2), (0.9, 1.0, 1.0), (0.9, 0.5, 0.5))
EXAMPLE_SELLERS = 2


class Coords:
    """(episode, step) coordinates of one env's contract streams; driven like a StepStream
    by harness.EpisodeClock."""

    def __init__(self, seed: int, env: int):
        self.seed, self.env = seed, env
        self.episode = self.step = 0
        self.reset_draws = 0

    def begin(self, episode: int, step: int) -> None:
        self.episode, self.step, self.reset_draws = episode, step, 0

    def d24(self, stream: int, idx: int) -> int:
        return rng.d24(self.seed, self.env, self.episode, self.step, stream, idx)


@contextlib.contextmanager
def contract_rng(coords: Coords, buyer_ordinal):
    """Route the example's three RNG call sites to the contract while the block runs.  The
    calling agent is recovered from the caller's frame (`self`), because the example calls
    module-level functions."""
    o_uniform, o_binomial, o_choice = np.random.uniform, np.random.binomial, random.choice

    def caller_ordinal():
        agent = sys._getframe(2).f_locals["self"]
        return buyer_ordinal[agent.id]

    def uniform(low=0.0, high=1.0, size=None):
        # reset-time draws are sequential in agent order; constructor-time draws (dead values,
        # overwritten by the first reset) use the same stream
        u = coords.d24(STREAM_RESET, coords.reset_draws) / 16777216.0
        coords.reset_draws += 1
        return np.asarray(low, np.float64) + (np.asarray(high, np.float64) - np.asarray(low, np.float64)) * u

    def binomial(n, p, size=None):
        assert n == 1
        return 1 if coords.d24(STREAM_DEMAND, caller_ordinal()) / 16777216.0 < p else 0

    def choice(seq):
        return seq[rng.randint(len(seq), coords.d24(STREAM_CHOICE, caller_ordinal()))]

    np.random.uniform, np.random.binomial, random.choice = uniform, binomial, choice
    try:
        yield
    finally:
        np.random.uniform, np.random.binomial, random.choice = o_uniform, o_binomial, o_choice


def _ids(n_buyers: int, n_sellers: int):
    return [f"b{i + 1}" for i in range(n_buyers)], [f"s{i + 1}" for i in range(n_sellers)]


def seller_handler_avg_price(env):
    """Env handler of the Sellers stage used by the handler-driven fixture (fsm.py:294-302): the
    sellers price again until the average price reaches 0.5.  Device twin:
    StageRule("Buyers", ("env", 1), ">=", 0x3FE00000, otherwise="Sellers")."""
    env.resolve_network()
    return "Buyers" if env.avg_price >= 0.5 else "Sellers"


def build_reference(buyers=EXAMPLE_BUYERS, n_sellers: int = EXAMPLE_SELLERS, num_steps: int = 10,
                    seller_stage_handler=None):
    """The unmodified example classes, wired the way example_simple_market.py:9-30 does.
    Returns (env, buyer_ordinal).  `seller_stage_handler`: the example's env object is built as
    is and its Sellers stage is then re-registered with that env handler and both next stages
    (the classes stay unmodified; only the stage table of this instance changes)."""
    from .. import ref_shim

    ph = ref_shim.import_reference()
    here = os.path.join(ref_shim.REFERENCE_ROOT, "examples/environments/simple_market")
    sys.path.insert(0, here)
    try:
        import market_agents  # noqa: the example's own module names
        import simple_mkt_env
    finally:
        sys.path.remove(here)
    Sampler = ph.utils.samplers.UniformFloatSampler
    buyer_ids, seller_ids = _ids(len(buyers), n_sellers)
    agents = [market_agents.BuyerAgent(b, p, supertype=market_agents.BuyerSupertype(Sampler(lo, hi)))
              for b, (p, lo, hi) in zip(buyer_ids, buyers)]
    agents += [market_agents.SellerAgent(s) for s in seller_ids]
    network = ph.Network(agents)
    network.add_connections_between(buyer_ids, seller_ids)
    env = simple_mkt_env.SimpleMarketEnv(num_steps=num_steps, network=network)
    if seller_stage_handler is not None:
        env._stages["Sellers"] = ph.FSMStage(
            stage_id="Sellers", acting_agents=seller_ids, rewarded_agents=seller_ids,
            next_stages=["Buyers", "Sellers"], handler=seller_stage_handler)
    return env, {b: i for i, b in enumerate(buyer_ids)}


def build(ph, UniformFloatSampler, buyers=EXAMPLE_BUYERS, n_sellers: int = EXAMPLE_SELLERS,
          num_steps: int = 10, seller_stage_handler=None):
    """The same env restated against plugin-API module `ph`.  Returns (env, buyer_ordinal)."""
    from ..phantom_oracle.spaces import Box, Discrete

    @dataclasses.dataclass(frozen=True)
    class Price(ph.MsgPayload):
        price: float

    @dataclasses.dataclass(frozen=True)
    class Order(ph.MsgPayload):
        vol: int

    @dataclasses.dataclass
    class BuyerSupertype(ph.Supertype):
        value: float

    class BuyerAgent(ph.StrategicAgent):
        def __init__(self, agent_id, demand_prob, supertype):
            super().__init__(agent_id, supertype=supertype)
            self.seller_prices, self.demand_prob, self.current_reward = {}, demand_prob, 0
            self.action_space = Discrete(2)
            self.observation_space = Box(low=0, high=1, shape=(3,))

        def decode_action(self, ctx, action):
            best = min(self.seller_prices.values())
            if not action:
                return []
            cheapest = [k for k, v in self.seller_prices.items() if v == best]
            self.current_reward += -action * best + self.type.value
            return [(random.choice(cheapest), Order(action))]

        def encode_observation(self, ctx):
            best = min(self.seller_prices.values())
            return np.array([best, np.random.binomial(1, self.demand_prob), self.type.value])

        def compute_reward(self, ctx):
            r, self.current_reward = self.current_reward, 0
            return r

        @ph.agents.msg_handler(Price)
        def on_price(self, ctx, message):
            self.seller_prices[message.sender_id] = message.payload.price

        def reset(self):
            super().reset()
            self.seller_prices, self.current_reward = {}, 0

    class SellerAgent(ph.StrategicAgent):
        def __init__(self, agent_id):
            super().__init__(agent_id)
            self.current_price = self.current_revenue = self.current_tx = 0
            self.action_space = Box(low=0, high=1, shape=(1,))
            self.observation_space = Box(np.array([0, 0]), np.array([np.inf, 1]))

        def decode_action(self, ctx, action):
            self.current_price = action
            return [(nid, Price(action)) for nid in ctx.neighbour_ids]

        def encode_observation(self, ctx):
            obs = np.array([self.current_tx, ctx.env_view.avg_price])
            self.current_tx = 0
            return obs

        def compute_reward(self, ctx):
            r, self.current_revenue = self.current_revenue, 0
            return r

        def reset(self):
            self.current_price = self.action_space.sample()
            self.current_revenue = self.current_tx = 0

        @ph.agents.msg_handler(Order)
        def on_order(self, ctx, message):
            self.current_revenue += self.current_price * message.payload.vol
            self.current_tx += message.payload.vol

    class SimpleMarketEnv(ph.FiniteStateMachineEnv):
        @dataclasses.dataclass(frozen=True)
        class View(ph.fsm.FSMEnvView):
            avg_price: float

        def __init__(self, num_steps, network, buyer_ids, seller_ids):
            self.avg_price = 0.0
            self._seller_ids = seller_ids
            # seller_stage_handler (not in the example): a Python env handler of the Sellers stage
            # (fsm.py:294-302), the twin of a device StageRule
            super().__init__(num_steps, network, initial_stage="Sellers", stages=[
                ph.FSMStage(stage_id="Buyers", next_stages=["Sellers"], acting_agents=buyer_ids,
                            rewarded_agents=buyer_ids),
                ph.FSMStage(stage_id="Sellers", acting_agents=seller_ids, rewarded_agents=seller_ids,
                            next_stages=(["Buyers"] if seller_stage_handler is None
                                         else ["Buyers", "Sellers"]),
                            handler=seller_stage_handler)])

        def view(self, neighbour_id=None):
            return self.View(avg_price=self.avg_price, **super().view({}).__dict__)

        def post_message_resolution(self):
            super().post_message_resolution()
            self.avg_price = np.mean([self.agents[s].current_price for s in self._seller_ids])

    buyer_ids, seller_ids = _ids(len(buyers), n_sellers)
    agents = [BuyerAgent(b, p, supertype=BuyerSupertype(UniformFloatSampler(lo, hi)))
              for b, (p, lo, hi) in zip(buyer_ids, buyers)]
    agents += [SellerAgent(s) for s in seller_ids]
    network = ph.Network(agents)
    network.add_connections_between(buyer_ids, seller_ids)
    env = SimpleMarketEnv(num_steps, network, buyer_ids, seller_ids)
    return env, {b: i for i, b in enumerate(buyer_ids)}


def to_action(env):
    """harness.run_generic action conversion: sellers get the python float of the float32
    action (the example's SellerPolicy returns a float), buyers int(round(a))."""
    ids = env.strategic_agent_ids
    is_buyer = [type(env.agents[a]).__name__ == "BuyerAgent" for a in ids]
    return lambda s, a: int(round(float(a[0]))) if is_buyer[s] else float(a[0])


def state(env) -> np.ndarray:
    """float64 [n_agents + 1, 3]: buyers (current_reward, type.value, n prices heard), sellers
    (current_price, current_revenue, current_tx), last row (avg_price, 0, 0)."""
    rows = []
    for a in env.agents.values():
        if type(a).__name__ == "BuyerAgent":
            rows.append([float(a.current_reward), float(a.type.value), float(len(a.seller_prices))])
        else:
            rows.append([float(np.asarray(a.current_price).reshape(-1)[0]),
                         float(a.current_revenue), float(a.current_tx)])
    rows.append([float(env.avg_price), 0.0, 0.0])
    return np.array(rows, np.float64)


def actions_for(n_env: int, n_ep: int, T: int, n_buyers: int, n_sellers: int, seed: int):
    """f32 [n_env, n_ep, T, S, 1] + mask: seller prices are quantised to sixteenths in half of
    the envs (ties between sellers exercise random.choice); buyers buy with p = 0.7; a buyer's
    action is withheld now and then (sellers always act: the reference's np.mean over a mix of
    sampled arrays and floats is not defined)."""
    r = np.random.RandomState(seed)
    S = n_buyers + n_sellers
    a = r.uniform(0, 1, size=(n_env, n_ep, T, S, 1)).astype(np.float32)
    a[:, :, :, :n_buyers] = (a[:, :, :, :n_buyers] < 0.7)
    q = np.floor(a[::2, :, :, n_buyers:] * 4) / 4
    a[::2, :, :, n_buyers:] = q.astype(np.float32)
    m = np.ones((n_env, n_ep, T, S), np.uint8)
    m[:, :, :, :n_buyers] = r.uniform(size=(n_env, n_ep, T, n_buyers)) > 0.1
    return a, m
