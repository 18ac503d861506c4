"""Shared helpers of the golden-trace parity tests (C3 / C4 / C5 and friends)."""
import numpy as np

GENERIC_KEYS = ["obs_mask", "reward_mask", "term", "trunc", "all_done"]


def assert_oracle_trace_equal(tr, g, e):
    """Object-level oracle trace `tr` of env e == golden `g` (bit-exact, incl. float64 rewards)."""
    for k in ["reset_obs", "reset_mask", "obs", "reward", "state"] + GENERIC_KEYS:
        if k in g.files:
            assert np.array_equal(tr[k], g[k][e]), (e, k)


def assert_device_step_equal(out, g, ep, t, ctx=""):
    """BatchStep `out` (device tensors, all envs of the golden) == golden step (ep, t).
    Integers / masks bit-exact; obs exact (float32 both sides); rewards == float32(reference
    float64) where present -- the spec tolerance is 1e-5 relative, we are exact."""
    om = out.obs_mask.cpu().numpy()
    assert np.array_equal(om, g["obs_mask"][:, ep, t]), f"obs_mask {ctx}"
    obs = out.observations.cpu().numpy()
    want = g["obs"][:, ep, t]
    sel = om.astype(bool)
    assert np.array_equal(obs[sel], want[sel]), f"obs {ctx}"
    rm = out.reward_mask.cpu().numpy()
    assert np.array_equal(rm, g["reward_mask"][:, ep, t]), f"reward_mask {ctx}"
    rew = out.rewards.cpu().numpy()
    sel = rm == 1
    assert np.array_equal(rew[sel], g["reward"][:, ep, t].astype(np.float32)[sel]), f"reward {ctx}"
    assert np.array_equal(out.terminations.cpu().numpy(), g["term"][:, ep, t]), f"term {ctx}"
    assert np.array_equal(out.truncations.cpu().numpy(), g["trunc"][:, ep, t]), f"trunc {ctx}"
    assert np.array_equal(out.all_done.cpu().numpy(), g["all_done"][:, ep, t]), f"all_done {ctx}"


def device_trace_rows(env, e=0):
    counts, rows = env.tracked_messages_batch(e, e + 1)
    r = rows[0, : counts[0]]
    return np.stack([r[:, 0] & 0xFF, (r[:, 0] >> 8) & 0xFF, (r[:, 0] >> 16) & 0xFF, r[:, 1], r[:, 2]], 1)


def assert_batchsteps_equal(a, b, ctx=""):
    """Two BatchSteps agree: masks / flags bit-exact, obs and rewards wherever their mask says a
    value is present (rows with mask 0 are left untouched by the C ABI and may hold anything)."""
    import torch

    for name in ("obs_mask", "reward_mask", "terminations", "truncations", "all_done"):
        assert torch.equal(getattr(a, name), getattr(b, name)), f"{name} {ctx}"
    sel = a.obs_mask.bool()
    assert torch.equal(a.observations[sel], b.observations[sel]), f"observations {ctx}"
    sel = a.reward_mask == 1
    assert torch.equal(a.rewards[sel], b.rewards[sel]), f"rewards {ctx}"


def run_device_vs_golden(make_env, g, state_fn=None, state_every=1):
    """Step a device env (one handle holding every env of the golden) through the golden's
    action tape and compare every step."""
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    n_env, n_ep, T = A.shape[:3]
    env = make_env(num_envs=n_env, seed=seed)
    for ep in range(n_ep):
        obs, mask = env.reset_batch()
        assert np.array_equal(mask.cpu().numpy(), g["reset_mask"][:, ep]), f"reset mask ep {ep}"
        sel = g["reset_mask"][:, ep].astype(bool)
        assert np.array_equal(obs.cpu().numpy()[sel], g["reset_obs"][:, ep][sel]), f"reset obs {ep}"
        for t in range(T):
            out = env.step_batch(A[:, ep, t], M[:, ep, t])
            assert_device_step_equal(out, g, ep, t, f"ep {ep} t {t}")
            if state_fn is not None and (t % state_every == 0 or t == T - 1):
                assert np.array_equal(state_fn(env), g["state"][:, ep, t]), (ep, t)
    env.check_errors()
    return env


def run_device_trace_vs_golden(make_env, g, n_env):
    """Device message trace == the reference's Resolver.tracked_messages, env by env."""
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    gm = g["messages"]  # (env, ep, t, sender, recv, type, v0, v1)
    env = make_env(num_envs=n_env, seed=seed, enable_tracking=True)
    for ep in range(A.shape[1]):
        env.reset_batch()
        for t in range(A.shape[2]):
            env.step_batch(A[:n_env, ep, t], M[:n_env, ep, t])
            for e in range(n_env):
                want = gm[(gm[:, 0] == e) & (gm[:, 1] == ep) & (gm[:, 2] == t)][:, 3:8]
                assert np.array_equal(device_trace_rows(env, e), want), (e, ep, t)
    env.close()


def run_device_rollout_trace_vs_golden(make_env, g, n_env):
    """The message trace recorded INSIDE one T-step rollout launch (one slab per (step, env),
    phx_get_trace_step) == the reference's Resolver.tracked_messages of every step of a whole
    episode, env by env; the output planes of the tracked launch equal the golden too."""
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    gm = g["messages"]  # (env, ep, t, sender, recv, type, v0, v1)
    env = make_env(num_envs=n_env, seed=seed, enable_tracking=True)
    T = A.shape[2]
    for ep in range(A.shape[1]):
        env.reset_batch()
        acts = np.ascontiguousarray(np.swapaxes(A[:n_env, ep], 0, 1))   # [T, E, S, A]
        mask = np.ascontiguousarray(np.swapaxes(M[:n_env, ep], 0, 1))   # [T, E, S]
        out = env.rollout_batch(acts, mask)
        env.check_errors()
        for t in range(T):
            counts, rows = env.tracked_messages_batch(0, n_env, step=t)
            for e in range(n_env):
                r = rows[e, : counts[e]]
                got = np.stack([r[:, 0] & 0xFF, (r[:, 0] >> 8) & 0xFF, (r[:, 0] >> 16) & 0xFF,
                                r[:, 1], r[:, 2]], 1)
                want = gm[(gm[:, 0] == e) & (gm[:, 1] == ep) & (gm[:, 2] == t)][:, 3:8]
                assert np.array_equal(got, want), (e, ep, t)
        om = out.obs_mask.cpu().numpy()
        assert np.array_equal(np.swapaxes(om, 0, 1), g["obs_mask"][:n_env, ep])
    env.close()
