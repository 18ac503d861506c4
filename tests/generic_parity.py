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
