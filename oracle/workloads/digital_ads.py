"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The reference's third example environment, examples/environments/digital_ads_market/
digital_ads_market.py, imported UNMODIFIED and driven under the contract RNG (build container
only; the GPU box sees the fixture tests/golden/digital_ads_reference.npz):

  :140-193  PublisherAgent   publisher_step: ImpressionRequest(user = np.random.choice([1, 2]))
                             to the exchange; on Ads: clicked ~ binomial(1, p[user][theme]) ->
                             ImpressionResult to the winning advertiser
  :196-363  AdvertiserAgent  caches the user of the impression; advertiser_step: bid =
                             min(action * budget, left) -> Bid to the exchange; books the cost
                             of a won auction, counts clicks; obs {budget_left, type.budget,
                             user_id} (None before the first impression); reward = clicks of the
                             step; terminates when the budget is spent
  :366-515  AdExchangeAgent  forwards impressions to every advertiser; handle_batch OVERRIDE:
                             all Bids of the batch -> one first/second price auction (stable
                             descending sort: the first of equal highest bids wins) -> Ads to
                             the publisher, AuctionResult to every bidder
  :518-591  DigitalAdsEnv    two FSM stages, StochasticNetwork (all rates 1.0),
                             ignore_connection_errors, BatchResolver(round_limit=5)

RNG call sites -> contract (oracle/rng.py, 24-bit draws, `contract_rng`):
  stream 3   env-managed samplers at reset (UniformFloatSampler with clipping, samplers.py:142-147),
             step 0, idx = position in env._samplers
  stream 5   StochasticNetwork.resample_connectivity (network.py:446), as everywhere else
  stream 9   np.random.choice([1, 2]) of the impression (:52), idx 0
  stream 10  np.random.binomial(1, p) of the click (:190), idx 0
Observation layout in the C ABI (float32 [3]): [budget_left, type.budget, user_id - 1]
(gymnasium's Dict space orders its keys alphabetically: budget_left, type, user_id).
Device twin: phantom_b200/csrc/fam_digital_ads.cu.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os

import numpy as np

from .. import rng

STREAM_SAMPLER, STREAM_CONNECTIVITY, STREAM_USER, STREAM_CLICK = 3, 5, 9, 10
OBS_DIM = 3
THEMES = ("sport", "travel", "science", "tech")
# DigitalAdsEnv.__init__ :523-526 (the table the env hands to its publisher)
CLICK_PROBA = {1: {"sport": 0.0, "travel": 1.0, "science": 0.2, "tech": 0.5},
               2: {"sport": 1.0, "travel": 0.0, "science": 0.7, "tech": 0.5}}


class Coords:
    def __init__(self, seed: int, env: int):
        self.seed, self.env = seed, env
        self.episode = self.step = 0
        self.k = {}

    def begin(self, episode: int, step: int) -> None:
        self.episode, self.step, self.k = episode, step, {}

    def next(self, stream: int) -> int:
        i = self.k.get(stream, 0)
        self.k[stream] = i + 1
        return rng.d24(self.seed, self.env, self.episode, self.step, stream, i)


@contextlib.contextmanager
def contract_rng(coords: Coords):
    o = (np.random.uniform, np.random.binomial, np.random.choice, np.random.random)

    def uniform(low=0.0, high=1.0, size=None):
        return np.float64(rng.uniform_f64(low, high, coords.next(STREAM_SAMPLER)))

    def binomial(n, p, size=None):
        assert n == 1
        return 1 if coords.next(STREAM_CLICK) / 16777216.0 < p else 0

    def choice(seq, *a, **k):
        return seq[rng.randint(len(seq), coords.next(STREAM_USER))]

    def random(size=None):
        return coords.next(STREAM_CONNECTIVITY) / 16777216.0

    np.random.uniform, np.random.binomial, np.random.choice, np.random.random = uniform, binomial, choice, random
    try:
        yield
    finally:
        np.random.uniform, np.random.binomial, np.random.choice, np.random.random = o


def import_reference_module():
    from .. import ref_shim

    ref_shim.import_reference()
    path = os.path.join(ref_shim.REFERENCE_ROOT,
                        "examples/environments/digital_ads_market/digital_ads_market.py")
    spec = importlib.util.spec_from_file_location("_ref_digital_ads", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def build_reference(num_agents_theme, budgets, num_steps: int = 20, strategy: str = "first"):
    """budgets: per advertiser (low, high, clip_low, clip_high) of its UniformFloatSampler, in
    ADV_1.. order -- the `train` configuration of the example (:795-842) in small."""
    mod = import_reference_module()
    import phantom as ph  # the reference package (on sys.path once the shim is installed)

    Sampler = ph.utils.samplers.UniformFloatSampler
    supertypes = {f"ADV_{i + 1}": mod.AdvertiserAgent.Supertype(budget=Sampler(*b))
                  for i, b in enumerate(budgets)}
    env = mod.DigitalAdsEnv(num_steps=num_steps, num_agents_theme=dict(num_agents_theme),
                            agent_supertypes=supertypes)
    env.agents["ADX"].strategy = strategy
    return env


def flatten_obs(obs) -> np.ndarray:
    # type.to_obs_space_compatible_type() wraps floats as shape-(1,) arrays (supertype.py:63-89)
    budget = np.asarray(obs["type"]["budget"], np.float64).reshape(-1)[0]
    return np.array([obs["budget_left"][0], budget, obs["user_id"]], np.float32)


def state(env) -> np.ndarray:
    """float64 [n_advertisers, 8]: left, type.budget, bid, step_clicks, step_wins, current user,
    total_requests / total_wins of user 1."""
    rows = []
    for aid, a in env.agents.items():
        if not aid.startswith("ADV"):
            continue
        rows.append([float(a.left), float(a.type.budget), float(a.bid), a.step_clicks, a.step_wins,
                     float(a._current_user_id), a.total_requests[1], a.total_wins[1]])
    return np.array(rows, np.float64)


def actions_for(n_env, n_ep, T, S, seed):
    """bids as budget fractions; half of the envs quantise them to eighths (equal highest bids
    exercise the stable sort), and some are zero (no Bid message at all)."""
    r = np.random.RandomState(seed)
    a = r.uniform(0, 0.6, size=(n_env, n_ep, T, S, 1)).astype(np.float32)
    a[::2] = (np.floor(a[::2] * 8) / 8).astype(np.float32)
    a[r.uniform(size=a.shape) < 0.1] = 0.0
    m = (r.uniform(size=(n_env, n_ep, T, S)) > 0.05).astype(np.uint8)
    return a, m
