"""The reference's own test agents on the device (family PHX_FAMILY_MOCK, csrc/fam_mock.cu).

`MockAgent` / `MockStrategicAgent` mirror /root/reference/tests/__init__.py:28-69 (same
constructor keywords, same call counters), so the reference's step-loop known-answer tests
(tests/test_env.py, tests/fsm/*, tests/test_stackelberg.py) can be replayed verbatim against
the CUDA engine.  `EchoAgent` covers the message-passing agents of
tests/network/test_tracking.py and tests/network/test_resolver.py.
"""
from __future__ import annotations

from typing import Optional

import phantom_b200 as ph
from phantom_b200 import _lib as L
from phantom_b200.agents import device_column
from phantom_b200.families import FamilyInfo, register
from phantom_b200.spaces import Box

KIND_AGENT, KIND_STRATEGIC, KIND_ECHO, KIND_CODEC = 0, 1, 2, 3


@ph.msg_payload()
class TestMessage:
    __test__ = False  # not a pytest class
    value: int


@ph.msg_payload()
class Request:
    cash: int


@ph.msg_payload()
class Response:
    cash: int


class MockAgent(ph.Agent):
    __phx_family__ = "mock"
    __phx_kind__ = KIND_AGENT
    __phx_device_class__ = True

    def __init__(self, *args, num_steps: Optional[int] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.num_steps = num_steps


class MockStrategicAgent(ph.StrategicAgent):
    __phx_family__ = "mock"
    __phx_kind__ = KIND_STRATEGIC
    __phx_device_class__ = True

    encode_obs_count = device_column(0)
    decode_action_count = device_column(1)
    compute_reward_count = device_column(2)

    def __init__(self, *args, num_steps: Optional[int] = None, **kwargs):
        super().__init__(*args, **kwargs)
        self.action_space = Box(0, 1, (1,))
        self.observation_space = Box(0, 1, (1,))
        self.num_steps = num_steps


class CodecAgent(ph.StrategicAgent):
    """A StrategicAgent assembled from Encoder / Decoder / RewardFunction objects, with no
    behaviour of its own (reference: agents.py:199-290 -- the encoder / decoder / reward
    function are consulted by the default encode_observation / decode_action /
    compute_reward).  All three must be device-lowerable."""

    __phx_family__ = "mock"
    __phx_kind__ = KIND_CODEC
    __phx_device_class__ = True

    encode_obs_count = device_column(0)
    decode_action_count = device_column(1)
    compute_reward_count = device_column(2)

    def __init__(self, agent_id, observation_encoder, action_decoder, reward_function):
        super().__init__(agent_id, observation_encoder, action_decoder, reward_function)


class EchoAgent(ph.Agent):
    """seed_value > 0: every step, sends TestMessage / Request(seed_value) to each neighbour
    with a higher slot.  Handles TestMessage(v) by replying v // 2 while v > 1, Request(c) by
    replying Response(c // 2), Response by doing nothing."""

    __phx_family__ = "mock"
    __phx_kind__ = KIND_ECHO
    __phx_device_class__ = True

    handled_count = device_column(3)
    handled_total = device_column(4)
    level = device_column(5, dtype="float32")  # level * 0.5 + value per handled message (float32)

    def __init__(self, agent_id, seed_value: int = 0, request_response: bool = False):
        super().__init__(agent_id)
        self.seed_value = seed_value
        self.request_response = request_response


def _collect(env, agents, spec) -> None:
    from phantom_b200.errors import NotLowerableError

    for i, a in enumerate(agents):
        if isinstance(a, CodecAgent):
            for name in ("observation_encoder", "action_decoder", "reward_function"):
                if getattr(a, name) is None:  # agents.py:240-243,265-268,285-288
                    raise NotImplementedError(
                        f"Agent '{a.id}' does not have a {name} instance set")
            ops = a.observation_encoder.device_ops()
            if len(ops) > L.PHX_MAX_CODEC_OPS or sum(n for _, n, _ in ops) > 8:
                raise NotLowerableError(f"encoder of agent '{a.id}' is too large for the device")
            for k, (code, n, value) in enumerate(ops):
                spec.agent_codec_op[i][k] = code | (n << 8)
                spec.agent_codec_val[i][k] = value
            if a.action_decoder.flat_dim() > 1:
                pass  # wider Tuple / Dict action rows: only the declared width matters here
            a.action_decoder.device_ops()  # raises NotLowerableError for non-device decoders
            spec.agent_fparam[i][0] = a.reward_function.device_op()[1]
    for i, a in enumerate(agents):
        ns = getattr(a, "num_steps", None)
        spec.agent_iparam[i][0] = -1 if ns is None else int(ns)
        spec.agent_iparam[i][1] = int(getattr(a, "seed_value", 0))
        spec.agent_iparam[i][2] = int(bool(getattr(a, "request_response", False)))


FAMILY = register(FamilyInfo(
    name="mock",
    family_id=L.FAMILY_MOCK,
    payload_types=(TestMessage, Request, Response),
    obs_dim=8,
    act_dim=1,
    env_kinds=(L.ENV_BASE, L.ENV_FSM, L.ENV_STACKELBERG),
    collect=_collect,
    trace_capacity=lambda env, agents: 64,
))
