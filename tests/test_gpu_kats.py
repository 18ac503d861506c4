"""The reference's KATs (restated in tests/kat_scenarios.py) replayed through the CUDA engine:
phantom_b200's plugin API -> lowering -> C ABI -> generic queue engine, one env per handle."""
import pytest

pytestmark = pytest.mark.gpu

from . import kat_scenarios as kats  # noqa: E402


@pytest.fixture(scope="module")
def K():
    import phantom_b200 as ph
    from phantom_b200.envs import mock

    class NS:
        pass

    ns = NS()
    ns.ph = ph
    ns.MockAgent, ns.MockStrategicAgent, ns.EchoAgent = mock.MockAgent, mock.MockStrategicAgent, mock.EchoAgent
    ns.finish_network = lambda network: network
    ns.CodecAgent = mock.CodecAgent
    ns.ElapsedTime, ns.CurrentStep = ph.encoders.ElapsedTime, ph.encoders.CurrentStep
    return ns


@pytest.mark.parametrize("exec_mode", ["thread", "queue"])
@pytest.mark.parametrize("scenario", kats.ALL, ids=lambda f: f.__name__)
def test_kat_on_device(K, scenario, exec_mode, monkeypatch):
    """Every KAT on both variants of the generic engine: one thread per env, and a tile of
    lanes per env."""
    monkeypatch.setattr(K.ph.PhantomEnv, "default_exec_mode", exec_mode)
    env = scenario(K)
    if env is not None:
        assert env.exec_name.startswith("thread-per-env" if exec_mode == "thread" else "queue(G=")
        env.close()


def test_batched_kat_every_env_identical(K):
    """The same FSM KAT for 3 000 envs at once: every env must reproduce the single-env trace
    (mock agents draw no random numbers)."""
    import numpy as np

    ph = K.ph
    network = ph.Network([K.MockStrategicAgent("odd_agent"), K.MockStrategicAgent("even_agent")])
    env = ph.FiniteStateMachineEnv(
        num_steps=3, network=network, initial_stage="ODD", num_envs=3000,
        stages=[
            ph.FSMStage(stage_id="ODD", next_stages=["EVEN"], acting_agents=["odd_agent"],
                        rewarded_agents=["odd_agent"]),
            ph.FSMStage(stage_id="EVEN", next_stages=["ODD"], acting_agents=["even_agent"],
                        rewarded_agents=["even_agent"]),
        ])
    obs, mask = env.reset_batch()
    assert mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    acts = np.zeros((3000, 2, 1), np.float32)
    out = env.step_batch(acts, np.array([[1, 0]] * 3000, np.uint8))
    assert out.obs_mask.cpu().numpy().tolist() == [[0, 1]] * 3000
    assert out.reward_mask.cpu().numpy().tolist() == [[0, 2]] * 3000  # even_agent: None
    np.testing.assert_allclose(out.observations.cpu().numpy()[:, 1, 0], 1 / 3, rtol=1e-6)
    out = env.step_batch(acts, np.array([[0, 1]] * 3000, np.uint8))
    assert out.obs_mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    assert out.reward_mask.cpu().numpy().tolist() == [[1, 0]] * 3000
    out = env.step_batch(acts, np.array([[1, 0]] * 3000, np.uint8))
    assert out.all_done.cpu().numpy().tolist() == [[0, 1]] * 3000
    assert out.obs_mask.cpu().numpy().tolist() == [[1, 1]] * 3000  # terminal flush of the caches
    assert out.reward_mask.cpu().numpy().tolist() == [[1, 1]] * 3000
    env.close()
