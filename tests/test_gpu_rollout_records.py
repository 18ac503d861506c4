"""message tracing INSIDE a T-step rollout launch and the reference's
Rollout / Step / AgentStep records + JSONL export built from one device rollout
(phantom/utils/rollout.py:23-57,302-341; phantom/resolvers.py:41-60)."""
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from .generic_parity import run_device_rollout_trace_vs_golden  # noqa: E402


def test_rollout_trace_stackelberg_thread_and_tile(golden_dir):
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    g = np.load(os.path.join(golden_dir, "stackelberg_reference.npz"))
    for mode in ("thread", "queue"):
        run_device_rollout_trace_vs_golden(lambda **kw: StackelbergGameEnv(exec_mode=mode, **kw), g, 4)


def test_rollout_trace_market_whole_episode(golden_dir):
    """9 782 messages of env 0 over two 99-step episodes, each episode ONE launch."""
    from phantom_b200.envs.market import MarketEnv

    g = np.load(os.path.join(golden_dir, "market_reference.npz"))
    run_device_rollout_trace_vs_golden(lambda **kw: MarketEnv(**kw), g, 1)


def test_rollout_trace_dense_sparse_graph(golden_dir):
    from phantom_b200.envs.dense import DenseEnv

    g = np.load(os.path.join(golden_dir, "dense12_reference.npz"))
    run_device_rollout_trace_vs_golden(lambda **kw: DenseEnv(12, g["adjacency"], **kw), g, 6)


def test_rollout_trace_supply_chain_fast_kernel(golden_dir):
    """The schedule-specialised kernel records the trace of every step of a 100-step episode in
    one launch (golden: the reference's own supply_chain.py under the contract RNG)."""
    from phantom_b200.envs import supply_chain as sc

    g = np.load(os.path.join(golden_dir, "supply_chain_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    gm = g["messages"]  # (env, ep, t, sender, recv, type, size)
    n_env = int(gm[:, 0].max()) + 1
    env = sc.SupplyChainEnv(num_envs=n_env, seed=seed, enable_tracking=True)
    assert env.exec_name.startswith("fast")
    for ep in range(A.shape[1]):
        env.reset_batch()
        acts = np.ascontiguousarray(np.swapaxes(A[:n_env, ep], 0, 1)).reshape(A.shape[2], n_env, 1, 1)
        mask = np.ascontiguousarray(np.swapaxes(M[:n_env, ep], 0, 1)).reshape(A.shape[2], n_env, 1)
        env.rollout_batch(acts, mask)
        for t in range(A.shape[2]):
            counts, rows = env.tracked_messages_batch(0, n_env, step=t)
            for e in range(n_env):
                r = rows[e, : counts[e]]
                got = np.stack([r[:, 0] & 0xFF, (r[:, 0] >> 8) & 0xFF, (r[:, 0] >> 16) & 0xFF, r[:, 1]], 1)
                want = gm[(gm[:, 0] == e) & (gm[:, 1] == ep) & (gm[:, 2] == t)][:, 3:7]
                assert np.array_equal(got, want), (e, ep, t)
    env.check_errors()
    env.close()


def test_rollout_records_and_jsonl_roundtrip():
    """rollouts_from_batch: per-env Rollout records of one launch == stepping a single env
    through the dict API; rollouts_to_jsonl writes one parseable document per rollout with the
    reference's field names, messages included."""
    import phantom_b200 as ph
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv
    from phantom_b200.utils.rollout import (AgentStep, Rollout, Step, rollouts_from_batch,
                                            rollouts_to_dataframe, rollouts_to_jsonl)

    E, T, seed = 5, 20, 11
    r = np.random.RandomState(2)
    A = r.uniform(0, 1, size=(T, E, 4, 1)).astype(np.float32)
    M = (r.uniform(size=(T, E, 4)) > 0.2).astype(np.uint8)
    env = StackelbergGameEnv(num_envs=E, seed=seed, num_steps=T, enable_tracking=True)
    env.reset_batch()
    out = env.rollout_batch(A, M)
    revenue = np.asarray(env.agents["LEADER"].revenue_round)
    rollouts = rollouts_from_batch(env, A, out, M, record_messages=True,
                                   env_config={"n_followers": 3}, rollout_params={"seed": seed},
                                   metrics={"revenue": revenue})
    assert len(rollouts) == E and all(isinstance(x, Rollout) for x in rollouts)
    # env 3 against a single env object stepped through the reference's dict API
    e = 3
    one = StackelbergGameEnv(seed=seed, num_steps=T, env_offset=e, enable_tracking=True)
    one.reset()
    ids = one.strategic_agent_ids
    for t in range(T):
        acts = {aid: A[t, e, s] for s, aid in enumerate(ids) if M[t, e, s]}
        one.network.resolver.clear_tracked_messages()
        want = one.step(acts)
        got = rollouts[e][t]
        assert isinstance(got, Step) and got.i == t
        assert set(got.observations) == set(want.observations)
        for aid in want.observations:
            assert np.array_equal(got.observations[aid], want.observations[aid])
        assert got.rewards == want.rewards
        assert got.terminations == want.terminations and got.truncations == want.truncations
        assert set(got.actions) == set(acts)
        assert got.messages == one.network.resolver.tracked_messages
    agent_steps = rollouts[e].steps_for_agent("LEADER")
    assert isinstance(agent_steps[0], AgentStep) and len(agent_steps) == T
    assert rollouts[e].rewards_for_agent("F1", drop_nones=True) == [
        s.rewards["F1"] for s in rollouts[e].steps if s.rewards.get("F1") is not None]
    assert rollouts[e].metrics["revenue"] == revenue[e]
    buf = io.StringIO()
    rollouts_to_jsonl(rollouts, buf)
    lines = buf.getvalue().strip().split("\n")
    assert len(lines) == E
    doc = json.loads(lines[e])
    assert doc["rollout_id"] == e and doc["rollout_params"] == {"seed": seed}
    assert len(doc["steps"]) == T and set(doc["steps"][0]) == {
        "i", "observations", "rewards", "terminations", "truncations", "infos", "actions",
        "messages", "stage"}
    n_msgs = sum(len(s["messages"]) for s in doc["steps"])
    assert n_msgs == sum(len(s.messages) for s in rollouts[e].steps) > 0
    assert doc["steps"][1]["messages"][0].keys() >= {"sender_id", "receiver_id", "payload"}
    df = rollouts_to_dataframe(rollouts)
    assert float(df.loc[seed, "revenue"]) == pytest.approx(float(revenue.mean()))
    env.close(); one.close()
