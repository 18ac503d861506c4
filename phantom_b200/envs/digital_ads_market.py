"""The reference's third example environment on the device (family PHX_FAMILY_DIGITAL_ADS,
csrc/fam_digital_ads.cu): examples/environments/digital_ads_market/digital_ads_market.py -- a
publisher offers impressions, advertisers bid for them through an exchange that runs the auction
inside a `handle_batch` override, the publisher reports clicks.  Same class names, constructor
signatures, payload classes and env wiring as the example (:28-591), so its `env_config`
(`num_agents_theme`, `agent_supertypes` with clipped UniformFloatSamplers) builds this env.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Iterable

import numpy as np

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.errors import NotLowerableError
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box
from phantom_b200.utils.samplers import KIND_UNIFORM_FLOAT, Sampler

KIND_EXCHANGE, KIND_PUBLISHER, KIND_ADVERTISER = 0, 1, 2
THEMES = ("sport", "travel", "science", "tech")


@ph.msg_payload()
class ImpressionRequest:
    user_id: int


@ph.msg_payload()
class Bid:
    bid: float

    @classmethod
    def _phx_decode(cls, lo: int, hi: int) -> "Bid":
        return cls(_f64(lo, hi))


@ph.msg_payload()
class AuctionResult:
    cost: float

    @classmethod
    def _phx_decode(cls, lo: int, hi: int) -> "AuctionResult":
        return cls(_f64(lo, hi))


@ph.msg_payload()
class Ads:
    advertiser_slot: int
    theme_user: int  # theme id | user id << 8


@ph.msg_payload()
class ImpressionResult:
    clicked: int


def _f64(lo: int, hi: int) -> float:
    bits = (np.uint64(hi & 0xFFFFFFFF) << np.uint64(32)) | np.uint64(lo & 0xFFFFFFFF)
    return float(np.array([bits], np.uint64).view(np.float64)[0])


def _f64_column(env, agent, word):
    lo = env.agent_column(agent, word).astype(np.uint32).astype(np.uint64)
    hi = env.agent_column(agent, word + 1).astype(np.uint32).astype(np.uint64)
    v = ((hi << np.uint64(32)) | lo).view(np.float64)
    return v.item() if v.size == 1 else v


class PublisherAgent(ph.Agent):
    __phx_family__ = "digital_ads_market"
    __phx_kind__ = KIND_PUBLISHER
    __phx_device_class__ = True

    _USER_CLICK_PROBABILITIES = {
        1: {"sport": 0.0, "travel": 1.0, "science": 0.2, "tech": 0.8},
        2: {"sport": 1.0, "travel": 0.0, "science": 0.7, "tech": 0.1},
    }

    def __init__(self, agent_id: str, exchange_id: str, user_click_proba: dict = None):
        super().__init__(agent_id)
        self.exchange_id = exchange_id
        self.user_click_proba = user_click_proba or self._USER_CLICK_PROBABILITIES


class AdvertiserAgent(ph.StrategicAgent):
    """action Box(0, 1): the bid as a fraction of the budget; obs [budget_left, type.budget,
    user_id] (the example's Dict observation in gymnasium's key order); reward = clicks of the step."""

    __phx_family__ = "digital_ads_market"
    __phx_kind__ = KIND_ADVERTISER
    __phx_device_class__ = True

    @dataclasses.dataclass
    class Supertype(ph.Supertype):
        budget: float

    def __init__(self, agent_id: str, exchange_id: str, theme: str = "generic"):
        self.exchange_id = exchange_id
        self.theme = theme
        self.action_space = Box(low=np.array([0.0]), high=np.array([1.0]))
        super().__init__(agent_id)
        self.observation_space = Box(low=0.0, high=np.inf, shape=(3,))

    left = property(lambda self: _f64_column(self._phx_env, self, 0))
    bid = property(lambda self: _f64_column(self._phx_env, self, 4))
    step_clicks = property(lambda self: self._phx_env.agent_column(self, 6))
    step_wins = property(lambda self: self._phx_env.agent_column(self, 7))
    _current_user_id = property(lambda self: self._phx_env.agent_column(self, 8))

    @property
    def type(self):
        return self.Supertype(budget=_f64_column(self._phx_env, self, 2))

    def _per_user(self, word):
        env = self._phx_env
        return {1: env.agent_column(self, word), 2: env.agent_column(self, word + 1)}

    total_requests = property(lambda self: self._per_user(9))
    total_wins = property(lambda self: self._per_user(11))
    total_clicks = property(lambda self: self._per_user(13))


class AdExchangeAgent(ph.Agent):
    __phx_family__ = "digital_ads_market"
    __phx_kind__ = KIND_EXCHANGE
    __phx_device_class__ = True

    def __init__(self, agent_id: str, publisher_id: str, advertiser_ids: Iterable = tuple(),
                 strategy: str = "first"):
        super().__init__(agent_id)
        self.publisher_id = publisher_id
        self.advertiser_ids = advertiser_ids
        self.strategy = strategy


def _collect(env, agents, spec) -> None:
    exch = [a for a in agents if isinstance(a, AdExchangeAgent)]
    pubs = [a for a in agents if isinstance(a, PublisherAgent)]
    advs = [a for a in agents if isinstance(a, AdvertiserAgent)]
    if len(exch) != 1 or len(pubs) != 1 or not advs:
        raise NotLowerableError("digital-ads device program: one exchange, one publisher, >= 1 advertiser")
    ex, pub = exch[0], pubs[0]
    if list(ex.advertiser_ids) != [a.id for a in advs] or ex.publisher_id != pub.id:
        raise NotLowerableError("the exchange must list every advertiser, in agent order")
    if ex.strategy not in ("first", "second"):
        raise ValueError(f"Unknown auction strategy: {ex.strategy}")
    spec.iparams[0], spec.iparams[1] = ex._phx_slot, pub._phx_slot
    spec.iparams[2] = int(ex.strategy == "second")
    for u in (1, 2):
        for t, theme in enumerate(THEMES):
            p = float(pub.user_click_proba[u][theme])
            spec.iparams[3 + (u - 1) * 4 + t] = int(math.ceil(p * 16777216.0))
    samplers = getattr(env, "_samplers", [])
    for a in advs:
        i = a._phx_slot
        if a.exchange_id != ex.id or a.theme not in THEMES:
            raise NotLowerableError(f"advertiser '{a.id}': unknown exchange or theme '{a.theme}'")
        spec.agent_iparam[i][1] = THEMES.index(a.theme)
        if a.supertype is None:
            raise NotLowerableError(f"advertiser '{a.id}' needs a Supertype(budget=...)")
        b = a.supertype.budget
        if not isinstance(b, Sampler) or not any(b is s for s in samplers):
            raise NotLowerableError(
                "digital-ads device program: budgets are env-managed UniformFloatSamplers (the "
                "example's `agent_supertypes`); a constant python float budget makes the reference "
                "compute bids in float32, which is not restated here")
        kind, low, high = b.device_desc(allow_clip=True)
        if kind != KIND_UNIFORM_FLOAT:
            raise NotLowerableError("budget: only UniformFloatSampler is lowered")
        spec.agent_iparam[i][0] = next(k for k, s in enumerate(samplers) if s is b)
        spec.agent_fparam[i][0], spec.agent_fparam[i][1] = low, high
        big = 1.7976931348623157e308
        spec.agent_fparam[i][2] = -big if b.clip_low is None else float(b.clip_low)
        spec.agent_fparam[i][3] = big if b.clip_high is None else float(b.clip_high)


FAMILY = register(FamilyInfo(
    name="digital_ads_market",
    family_id=L.FAMILY_DIGITAL_ADS,
    payload_types=(ImpressionRequest, Bid, AuctionResult, Ads, ImpressionResult),
    obs_dim=3,
    act_dim=1,
    env_kinds=(L.ENV_FSM,),
    collect=_collect,
    trace_capacity=lambda env, agents: 4 * len(agents),
    supports_supertypes=True,
))


class DigitalAdsEnv(ph.FiniteStateMachineEnv):
    def __init__(self, num_steps=20, num_agents_theme=None, strategy: str = "first", **kwargs):
        self.exchange_id, self.publisher_id = "ADX", "PUB"
        click = {1: {"sport": 0.0, "travel": 1.0, "science": 0.2, "tech": 0.5},
                 2: {"sport": 1.0, "travel": 0.0, "science": 0.7, "tech": 0.5}}
        publisher = PublisherAgent(self.publisher_id, exchange_id=self.exchange_id,
                                   user_click_proba=click)
        advertisers, i = [], 1
        for theme, n in num_agents_theme.items():
            for _ in range(n):
                advertisers.append(AdvertiserAgent(f"ADV_{i}", self.exchange_id, theme=theme))
                i += 1
        self.advertiser_ids = [a.id for a in advertisers]
        exchange = AdExchangeAgent(self.exchange_id, publisher_id=self.publisher_id,
                                   advertiser_ids=self.advertiser_ids, strategy=strategy)
        network = ph.StochasticNetwork([exchange, publisher] + advertisers,
                                       ph.resolvers.BatchResolver(round_limit=5),
                                       ignore_connection_errors=True)
        network.add_connections_between([self.exchange_id], [self.publisher_id])
        network.add_connections_between([self.exchange_id], self.advertiser_ids)
        network.add_connections_between([self.publisher_id], self.advertiser_ids)
        super().__init__(
            num_steps=num_steps, network=network, initial_stage="publisher_step",
            stages=[
                ph.FSMStage(stage_id="publisher_step", next_stages=["advertiser_step"],
                            acting_agents=[self.publisher_id], rewarded_agents=[self.publisher_id]),
                ph.FSMStage(stage_id="advertiser_step", next_stages=["publisher_step"],
                            acting_agents=self.advertiser_ids, rewarded_agents=self.advertiser_ids),
            ], **kwargs)
