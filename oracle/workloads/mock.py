"""TEST INFRASTRUCTURE (CPU oracle) -- never imported by the product.

The agents of the reference's own tests, restated against the plugin API:
  MockAgent / MockStrategicAgent   /root/reference/tests/__init__.py:28-69
  EchoAgent                        tests/network/test_tracking.py:21-26 (halving replies) and
                                   tests/network/test_resolver.py:24-45 (request/response),
                                   plus a generate_messages() that replaces those tests'
                                   hand-made n.send() calls: TestMessage / Request(seed) to every
                                   neighbour with a higher slot, in slot order.
Device twin: phantom_b200/csrc/fam_mock.cu.
"""
from __future__ import annotations

from typing import Optional

import numpy as np


def build_classes(ph):
    @ph.msg_payload()
    class TestMessage:
        __test__ = False
        value: int

    @ph.msg_payload()
    class Request:
        cash: int

    @ph.msg_payload()
    class Response:
        cash: int

    class MockAgent(ph.Agent):
        def __init__(self, *args, num_steps: Optional[int] = None, **kwargs):
            super().__init__(*args, **kwargs)
            self.num_steps = num_steps

    class MockStrategicAgent(ph.StrategicAgent):
        def __init__(self, *args, num_steps: Optional[int] = None, **kwargs):
            super().__init__(*args, **kwargs)
            self.encode_obs_count = 0
            self.decode_action_count = 0
            self.compute_reward_count = 0
            self.num_steps = num_steps

        def encode_observation(self, ctx):
            self.encode_obs_count += 1
            return np.array([ctx.env_view.proportion_time_elapsed])

        def decode_action(self, ctx, action):
            self.decode_action_count += 1
            return []

        def compute_reward(self, ctx):
            self.compute_reward_count += 1
            return 0.0

        def is_terminated(self, ctx):
            return ctx.env_view.current_step == self.num_steps

        def is_truncated(self, ctx):
            return ctx.env_view.current_step == self.num_steps

    class EchoAgent(ph.Agent):
        def __init__(self, agent_id, seed_value: int = 0, request_response: bool = False):
            super().__init__(agent_id)
            self.seed_value = seed_value
            self.request_response = request_response
            self.handled_count = 0
            self.handled_total = 0
            self.level = np.float32(0.0)  # float32 state: level * 0.5 + value per handled message
            self._slots = None  # filled by the env builder: agent id -> slot

        def _bump(self, value):
            self.handled_count += 1
            self.handled_total += value
            self.level = np.float32(self.level * np.float32(0.5) + np.float32(value))

        def generate_messages(self, ctx):
            if self.seed_value <= 0:
                return []
            mine = self._slots[self.id]
            higher = sorted((self._slots[n], n) for n in ctx.neighbour_ids if self._slots[n] > mine)
            cls = Request if self.request_response else TestMessage
            return [(n, cls(self.seed_value)) for _, n in higher]

        @ph.agents.msg_handler(TestMessage)
        def on_test_message(self, ctx, message):
            self._bump(message.payload.value)
            if message.payload.value > 1:
                return [(message.sender_id, TestMessage(message.payload.value // 2))]

        @ph.agents.msg_handler(Request)
        def on_request(self, ctx, message):
            self._bump(message.payload.cash)
            return [(message.sender_id, Response(message.payload.cash // 2))]

        @ph.agents.msg_handler(Response)
        def on_response(self, ctx, message):
            self._bump(message.payload.cash)
            return []

    Box = __import__("oracle.phantom_oracle.spaces", fromlist=["Box"]).Box

    class ElapsedTime(ph.encoders.Encoder):
        @property
        def observation_space(self):
            return Box(0.0, 1.0, (1,))

        def encode(self, ctx):
            return np.array([ctx.env_view.proportion_time_elapsed])

    class CurrentStep(ph.encoders.Encoder):
        @property
        def observation_space(self):
            return Box(0.0, np.inf, (1,))

        def encode(self, ctx):
            return np.array([float(ctx.env_view.current_step)])

    class CodecAgent(ph.StrategicAgent):
        """Pure composition: everything comes from the encoder / decoder / reward objects."""

    class NS:
        pass

    ns = NS()
    ns.ph = ph
    ns.TestMessage, ns.Request, ns.Response = TestMessage, Request, Response
    ns.MockAgent, ns.MockStrategicAgent, ns.EchoAgent = MockAgent, MockStrategicAgent, EchoAgent
    ns.NetworkError = ph.network.NetworkError
    ns.CodecAgent, ns.ElapsedTime, ns.CurrentStep = CodecAgent, ElapsedTime, CurrentStep

    def finish_network(network):
        slots = {aid: i for i, aid in enumerate(network.agent_ids)}
        for a in network.agents.values():
            if isinstance(a, EchoAgent):
                a._slots = slots
        return network

    ns.finish_network = finish_network

    def stage_handler(then, lhs="always", cmp="==", rhs=0, otherwise=None, resolve_network=True,
                      also=None, elifs=None):
        """A Python env stage handler (reference: phantom/fsm.py:294-302) with the meaning of
        phantom_b200.fsm.StageRule: [env.resolve_network()]; an if / elif / else chain whose
        branches are conjunctions of comparisons between the clock, agent attributes and
        constants."""
        import operator

        ops = {"<": operator.lt, "<=": operator.le, "==": operator.eq, "!=": operator.ne,
               ">=": operator.ge, ">": operator.gt}
        branches = [([(lhs, cmp, rhs)] + ([] if also is None else [tuple(also)]), then)]
        for terms, stage in elifs or ():
            single = isinstance(terms, tuple) and len(terms) == 3 and terms[1] in ops
            branches.append(([terms] if single else list(terms), stage))

        def handler(env):
            if resolve_network:
                env.resolve_network()
            if lhs == "always":
                return then

            def value(x):
                if isinstance(x, tuple):
                    return getattr(env.agents[x[1]], x[2])
                return env.current_step if x == "step" else x

            for terms, stage in branches:
                if all(ops[c](value(l), value(r)) for l, c, r in terms):
                    return stage
            return otherwise

        return handler

    ns.stage_handler = stage_handler
    return ns
