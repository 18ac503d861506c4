"""A USER-WRITTEN device program, compiled at run time and loaded through phx_create_user: the
reference's plugin contract -- "subclass Agent and write handlers" (phantom/agents.py:48-60,
122-155) -- without rebuilding libphx.so.  The fixture env (tests/user_program/) exists twice:
agent classes bound to auction_game.cu, and the same env with Python handlers that runs on the
oracle port of the reference step loop; both must agree step by step."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_bidders,E,exec_name", [
    (3, 24, "thread-per-env(G=8, user program)"),
    (40, 6, "wide(G=128, user program)")])  # 41 agents: the 128-lane block engine
def test_user_program_matches_python_handlers(n_bidders, E, exec_name):
    import oracle.phantom_oracle as po
    from oracle import harness

    from .user_program import auction_env as ae

    T, n_ep, S = 40, 2, n_bidders
    r = np.random.RandomState(4)
    A = r.uniform(0, 1, size=(E, n_ep, T, S, 1)).astype(np.float32)
    M = (r.uniform(size=(E, n_ep, T, S)) > 0.15).astype(np.uint8)
    env = ae.build_device(n_bidders, num_envs=E, seed=1, num_steps=T)
    assert env.exec_name == exec_name
    traces = []
    for e in range(E):
        ref = ae.build_reference(po, T, n_bidders)
        traces.append(harness.run_generic(ref, harness.EpisodeClock([]), A[e], M[e], 3))
    for ep in range(n_ep):
        obs, mask = env.reset_batch()
        assert np.array_equal(mask.cpu().numpy(), np.stack([t["reset_mask"][ep] for t in traces]))
        assert np.array_equal(obs.cpu().numpy(), np.stack([t["reset_obs"][ep] for t in traces]))
        out = env.rollout_batch(np.ascontiguousarray(np.swapaxes(A[:, ep], 0, 1)),
                                np.ascontiguousarray(np.swapaxes(M[:, ep], 0, 1)))
        env.check_errors()
        want = {k: np.stack([t[k][ep] for t in traces], axis=1) for k in
                ("obs", "obs_mask", "reward", "reward_mask", "term", "trunc", "all_done")}
        om = out.obs_mask.cpu().numpy()
        assert np.array_equal(om, want["obs_mask"])
        sel = om.astype(bool)
        assert np.array_equal(out.observations.cpu().numpy()[sel], want["obs"][sel])
        rm = out.reward_mask.cpu().numpy()
        assert np.array_equal(rm, want["reward_mask"])
        assert np.array_equal(out.rewards.cpu().numpy()[rm == 1], want["reward"].astype(np.float32)[rm == 1])
        assert np.array_equal(out.terminations.cpu().numpy(), want["term"])
        assert np.array_equal(out.truncations.cpu().numpy(), want["trunc"])
        assert np.array_equal(out.all_done.cpu().numpy(), want["all_done"])
        assert (want["term"] == 1).any(), "bidders retire on this tape (done dropout exercised)"
    # device state columns behave like the Python attributes
    assert np.asarray(env.agents["B1"].wins).shape == (E,)
    env.close()


def test_user_program_dict_api_and_errors():
    import phantom_b200 as ph

    from .user_program import auction_env as ae

    env = ae.build_device(enable_tracking=True)
    obs, _ = env.reset()
    assert set(obs) == {"B1", "B2", "B3"}
    step = env.step({"B1": np.array([0.3], np.float32), "B3": np.array([0.9], np.float32)})
    assert step.rewards["B1"] == 0.75 and step.rewards["B3"] == 0.5 and step.rewards["B2"] == 1.0
    msgs = env.network.resolver.tracked_messages
    assert [m.sender_id for m in msgs] == ["B1", "B3", "BOOK", "BOOK"]
    assert (msgs[3].payload.rank, msgs[3].payload.best) == (2, 90)
    env.close()
    from phantom_b200 import _lib as L

    import ctypes as C
    spec = env.spec
    h = C.c_void_p()
    assert L.lib.phx_create(C.byref(spec), 4, 0, 0, 0, C.byref(h)) == L.PHX_ERR_INVALID
    assert L.lib.phx_create_user(C.byref(spec), b"/nonexistent.cubin", 4, 0, 0, 0, C.byref(h)) == L.PHX_ERR_CUDA
