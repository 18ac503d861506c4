"""Rollout / Step / AgentStep records and the JSONL / DataFrame exporters (CPU; reference:
phantom/utils/rollout.py:23-341).  The device -> records path is tests/test_gpu_rollout_records.py."""
import io
import json

import numpy as np

import phantom_b200 as ph
from phantom_b200.utils.rollout import (AgentStep, Rollout, Step, rollouts_to_dataframe,
                                        rollouts_to_jsonl)


@ph.msg_payload()
class Ping:
    n: int


def make_rollout(rid, scale):
    steps = []
    for i in range(4):
        obs = {"a": np.array([i * scale, 1.0], np.float32)}
        if i % 2 == 0:
            obs["b"] = np.array([2.0], np.float32)
        steps.append(Step(i, obs, {"a": float(i), "b": None if i == 1 else 0.5},
                          {"a": False, "b": i == 3, "__all__": False},
                          {"a": False, "b": False, "__all__": i == 3}, {k: {} for k in obs},
                          {"a": np.float32(i)} if i != 2 else {},
                          [ph.Message("a", "b", Ping(i))], "EVEN" if i % 2 == 0 else "ODD"))
    return Rollout(rid, 0, {"k": 1}, {"scale": scale}, steps, {"m": np.float64(rid * 10.0)})


def test_rollout_helpers_follow_the_reference():
    r = make_rollout(7, 2.0)
    assert r[2].i == 2
    assert len(r.observations_for_agent("b")) == 4 and r.observations_for_agent("b")[1] is None
    assert len(r.observations_for_agent("b", drop_nones=True)) == 2
    assert r.rewards_for_agent("b", drop_nones=True) == [0.5, 0.5, 0.5]
    assert r.rewards_for_agent("b") == [0.5, None, 0.5, 0.5]
    assert r.terminations_for_agent("b", stages=["ODD"]) == [False, True]
    assert r.actions_for_agent("a", drop_nones=True) == [np.float32(0), np.float32(1), np.float32(3)]
    steps = r.steps_for_agent("b", stages=["EVEN"])
    assert [s.i for s in steps] == [0, 2] and isinstance(steps[0], AgentStep)
    assert steps[0].observation is not None and steps[0].done is False and steps[0].stage == "EVEN"
    assert r.steps_for_agent("b")[3].done is True
    assert dict(r.count_agent_actions("a"))[None] == 1
    assert sum(n for _, n in r.count_actions()) == 3


def test_jsonl_and_dataframe_export():
    rollouts = [make_rollout(0, 1.0), make_rollout(1, 1.0), make_rollout(2, 3.0)]
    buf = io.StringIO()
    rollouts_to_jsonl(rollouts, buf)
    docs = [json.loads(ln) for ln in buf.getvalue().strip().split("\n")]
    assert [d["rollout_id"] for d in docs] == [0, 1, 2]
    step = docs[2]["steps"][1]
    assert step["observations"]["a"] == [3.0, 1.0] and step["rewards"]["b"] is None
    assert step["messages"] == [{"sender_id": "a", "receiver_id": "b", "payload": {"n": 1}}]
    assert step["stage"] == "ODD" and docs[2]["metrics"]["m"] == 20.0
    pretty = io.StringIO()
    rollouts_to_jsonl(rollouts[:1], pretty, human_readable=True)
    assert pretty.getvalue().count("\n") > 20
    df = rollouts_to_dataframe(rollouts)
    assert df.loc[1.0, "m"] == 5.0 and df.loc[3.0, "m"] == 20.0
    assert len(rollouts_to_dataframe(rollouts, avg_over_repeats=False)) == 3
