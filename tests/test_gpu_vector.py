"""Vector-env façade: per-agent tensors + lazy per-env Step dict views."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_vector_env_matches_golden_and_lazy_views(golden_dir):
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv
    from phantom_b200.vector import VectorEnv

    g = np.load(os.path.join(golden_dir, "stackelberg_reference.npz"))
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    n_env = A.shape[0]
    vec = VectorEnv(StackelbergGameEnv(num_envs=n_env, seed=seed))
    ids = vec.agent_ids
    assert ids == ["LEADER", "F1", "F2", "F3"]
    obs = vec.reset()
    assert list(obs) == ["LEADER"] and obs["LEADER"].shape == (n_env, 2)
    assert np.array_equal(obs["LEADER"].cpu().numpy(), g["reset_obs"][:, 0, 0])
    for t in range(30):
        acts = {aid: A[:, 0, t, s] for s, aid in enumerate(ids)}
        mask = {aid: M[:, 0, t, s] for s, aid in enumerate(ids)}
        step = vec.step(acts, mask)
        for s, aid in enumerate(ids):
            om = g["obs_mask"][:, 0, t, s].astype(bool)
            assert np.array_equal(step.obs_mask[aid].cpu().numpy(), om)
            assert np.array_equal(step.observations[aid].cpu().numpy()[om], g["obs"][:, 0, t, s][om])
            assert np.array_equal(step.reward_mask[aid].cpu().numpy(), g["reward_mask"][:, 0, t, s])
            assert np.array_equal(step.terminations[aid].cpu().numpy(), g["term"][:, 0, t, s])
        assert np.array_equal(step.all_truncated.cpu().numpy(), g["all_done"][:, 0, t, 1].astype(bool))
        # lazy dict view of one sub-env == the reference's Step for that env
        e = t % n_env
        view = step.env(e)
        assert set(view.observations) == {aid for s, aid in enumerate(ids) if g["obs_mask"][e, 0, t, s]}
        for s, aid in enumerate(ids):
            if g["reward_mask"][e, 0, t, s] == 1:
                assert view.rewards[aid] == pytest.approx(g["reward"][e, 0, t, s], rel=1e-5, abs=1e-7)
            else:
                assert aid not in view.rewards
        assert view.truncations["__all__"] == bool(g["all_done"][e, 0, t, 1])
    vec.close()


def test_vector_env_missing_agent_falls_back_to_generate_messages():
    from phantom_b200.envs.supply_chain import SupplyChainEnv
    from phantom_b200.vector import VectorEnv

    vec = VectorEnv(SupplyChainEnv(num_envs=128, seed=1))
    vec.reset()
    step = vec.step({})  # SHOP absent: no StockRequest anywhere (env.py:330-333)
    assert (np.asarray(vec.env.agents["SHOP"].delivered_stock) == 0).all()
    assert step.observations["SHOP"].shape == (128, 3)
    vec.close()
