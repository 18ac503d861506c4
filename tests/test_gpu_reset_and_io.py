"""a17 / boundary details on the device: masked reset (vector-env auto-reset building block),
auto-reset inside engine rollouts, state column write-back, host-buffer rollouts and odd batch
sizes, for every kernel variant."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import vectorised  # noqa: E402

from .generic_parity import assert_batchsteps_equal  # noqa: E402


@pytest.mark.parametrize("exec_mode", ["fast", "thread", "queue"])
def test_masked_reset_only_touches_masked_envs(exec_mode):
    """phx_reset(env_mask): masked envs restart (clock 0, episode+1, stock 0, sales kept --
    supply_chain.py:149-150), the others continue untouched; checked against two vectorised
    oracles stepped in lock step."""
    import torch

    from phantom_b200 import _lib as L
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    E, seed = 777, 13  # odd size: partial warps / tiles / blocks
    env = SupplyChainEnv(num_envs=E, seed=seed, exec_mode=exec_mode)
    ref = vectorised.SupplyChainVec(E, seed)
    env.reset_batch(); ref.reset()
    r = np.random.RandomState(0)
    for t in range(40):
        a = r.uniform(0, 100, size=(E, 1, 1)).astype(np.float32)
        out = env.step_batch(a)
        want = ref.step(a[:, 0, 0])
        assert np.array_equal(out.observations.cpu().numpy()[:, 0], want["obs"]), t
        if t % 7 == 6:
            mask = r.uniform(size=E) < 0.3
            obs, om = env.reset_batch(torch.as_tensor(mask))
            want_obs = ref.reset(mask)
            got = obs.cpu().numpy()[:, 0]
            assert np.array_equal(got[mask], want_obs[mask])
            assert np.array_equal(om.cpu().numpy()[:, 0].astype(bool), mask)
            assert np.array_equal(env.field(L.FIELD_STEP, np.int32), ref.step_no)
            assert np.array_equal(env.field(L.FIELD_EPISODE, np.int32), ref.episode)
    env.check_errors()
    env.close()


@pytest.mark.parametrize("exec_mode", ["thread", "queue"])
def test_engine_auto_reset_rollout_equals_manual_resets(exec_mode):
    """PHX_FLAG_AUTO_RESET inside a T-step engine rollout == stepping and resetting by hand
    (Stackelberg: reward cache cleared, leaders observe the reset, followers do not)."""
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    E, T, seed = 300, 25, 4
    A = np.random.RandomState(1).uniform(0, 1, size=(T, E, 4, 1)).astype(np.float32)
    auto = StackelbergGameEnv(num_envs=E, seed=seed, num_steps=10, auto_reset=True, exec_mode=exec_mode)
    manual = StackelbergGameEnv(num_envs=E, seed=seed, num_steps=10, exec_mode=exec_mode)
    auto.reset_batch(); manual.reset_batch()
    ro = auto.rollout_batch(A)
    for t in range(T):
        out = manual.step_batch(A[t])
        done = bool(out.all_done.cpu().numpy()[0, 1])
        rew_a, rm_a = ro.rewards[t].cpu().numpy(), ro.reward_mask[t].cpu().numpy()
        assert np.array_equal(rm_a, out.reward_mask.cpu().numpy()), t
        assert np.array_equal(rew_a[rm_a == 1], out.rewards.cpu().numpy()[rm_a == 1]), t
        assert np.array_equal(ro.all_done[t].cpu().numpy(), out.all_done.cpu().numpy()), t
        if done:
            obs, om = manual.reset_batch()
            assert np.array_equal(ro.obs_mask[t].cpu().numpy(), om.cpu().numpy()), t
            sel = om.cpu().numpy().astype(bool)
            assert np.array_equal(ro.observations[t].cpu().numpy()[sel], obs.cpu().numpy()[sel]), t
        else:
            sel = out.obs_mask.cpu().numpy().astype(bool)
            assert np.array_equal(ro.obs_mask[t].cpu().numpy().astype(bool), sel), t
            assert np.array_equal(ro.observations[t].cpu().numpy()[sel], out.observations.cpu().numpy()[sel]), t
    auto.close(); manual.close()


def test_set_field_roundtrip_and_effect():
    """phx_set_field: writing a state column changes the dynamics accordingly."""
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    for mode in ("fast", "thread"):
        env = SupplyChainEnv(num_envs=16, seed=2, exec_mode=mode)
        env.reset_batch()
        shop = env.agents["SHOP"]
        shop.stock = np.arange(16) * 5
        assert np.array_equal(np.asarray(shop.stock), np.arange(16) * 5)
        out = env.step_batch(np.zeros((16, 1, 1), np.float32))
        stock_after = np.asarray(shop.stock)
        assert np.array_equal(stock_after, np.arange(16) * 5 - np.asarray(shop.sales))
        assert np.array_equal(out.observations.cpu().numpy()[:, 0, 0], (stock_after / 100).astype(np.float32))
        env.close()


def test_rollout_host_for_engine_families():
    """phx_rollout_host (host buffers, unchunked path) on the generic engine == device path."""
    from phantom_b200 import BatchStep
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    import torch

    E, T, seed = 1500, 12, 8
    A = np.random.RandomState(3).uniform(0, 1, size=(T, E, 4, 1)).astype(np.float32)
    a, b = StackelbergGameEnv(num_envs=E, seed=seed), StackelbergGameEnv(num_envs=E, seed=seed)
    a.reset_batch(); b.reset_batch()
    dev = a.rollout_batch(A)
    host = b.rollout_host(A)
    hb = BatchStep(*[torch.as_tensor(host[k]).cuda() for k in
                     ("observations", "obs_mask", "rewards", "reward_mask", "terminations", "truncations", "all_done")])
    assert_batchsteps_equal(dev, hb)
    a.close(); b.close()


def test_rollout_host_pipeline_chunks_along_time():
    """phx_rollout_host's three-stream pipeline splits a long rollout into 8 chunks of the time
    axis (T * E >= 65 536, T >= 8): uneven chunk lengths (T = 13 -> 1 or 2 steps each), an FSM
    env class whose reward / observation caches and stage must carry across the chunk launches,
    auto-reset inside a chunk -- all identical to one device rollout."""
    from phantom_b200 import BatchStep
    from phantom_b200.envs import simple_market as sm
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    import torch

    keys = ("observations", "obs_mask", "rewards", "reward_mask", "terminations", "truncations", "all_done")
    for make, S, T, E in ((lambda: sm.example_env(num_envs=6000, seed=4, num_steps=5, auto_reset=True), 5, 13, 6000),
                          (lambda: SupplyChainEnv(num_envs=8192, seed=4, num_steps=7, auto_reset=True), 1, 29, 8192)):
        A = np.random.RandomState(5).uniform(0, 1, size=(T, E, S, 1)).astype(np.float32)
        a, b = make(), make()
        a.reset_batch(); b.reset_batch()
        dev = a.rollout_batch(A)
        host = b.rollout_host(A)
        hb = BatchStep(*[torch.as_tensor(host[k]).cuda() for k in keys])
        assert_batchsteps_equal(dev, hb)
        a.check_errors(); b.check_errors()
        a.close(); b.close()


def _host_planes(T, E):
    import torch

    pin = lambda shape, dt: torch.empty(shape, dtype=dt).pin_memory()
    return pin((T, E, 1, 3), torch.float32), pin((T, E, 1), torch.float32), pin((T, E, 2), torch.uint8)


@pytest.mark.parametrize("lo,hi,in_range", [(0.0, 100.0, True), (-40.0, 140.0, False),
                                            (-60000.0, 100.0, False)])
def test_rollout_host_wire_chunks_equal_device_planes(lo, hi, in_range, monkeypatch):
    """The lean host-buffer call of the supply chain sends its first time chunks across PCIe as
    ONE 32-bit word per env-step and expands them on host threads (csrc/phx_sc_wire.*) while the
    DMA engine copies the float planes of the other chunks (opt-in: PHX_WIRE_CHUNKS).  Whatever the
    split (some / all chunks as wire words, the adaptive controller, or none), the planes must equal the device
    path's bit for bit -- in distribution, with negative stock / sales (outside the wire fields:
    the wire rows are re-copied from the float planes the kernel also wrote), across auto-reset
    wraps and over three consecutive calls."""
    import torch

    from phantom_b200 import _lib as L
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    E, T, seed = 8192, 37, 14
    r = np.random.RandomState(7)
    tapes = [r.uniform(lo, hi, size=(T, E, 1, 1)).astype(np.float32) for _ in range(3)]
    dev = SupplyChainEnv(num_envs=E, seed=seed, num_steps=13, auto_reset=True)
    dev.reset_batch()
    want = [[x.cpu() for x in dev.rollout_batch(A)] for A in tapes]
    want = [(w[0], w[2], w[6]) for w in want]
    h_obs, h_rew, h_all = _host_planes(T, E)
    shop = lambda env: np.stack([np.asarray(getattr(env.agents["SHOP"], c))
                                 for c in ("stock", "sales", "missed_sales")], 1)
    if not in_range:  # the tape really leaves the wire fields
        assert shop(dev)[:, 0].min() < -32768 or shop(dev)[:, 1].min() < 0
    for chunks in ("5", "14", "auto", None):
        if chunks is None:  # the default: plain staged copies of the float planes
            monkeypatch.delenv("PHX_WIRE_CHUNKS", raising=False)
        else:
            monkeypatch.setenv("PHX_WIRE_CHUNKS", chunks)
        env = SupplyChainEnv(num_envs=E, seed=seed, num_steps=13, auto_reset=True)
        env.reset_batch()
        for A, (o, rw, ad) in zip(tapes, want):
            a = torch.as_tensor(A).pin_memory()
            h_obs.fill_(-1.0); h_rew.fill_(-1.0); h_all.fill_(7)
            L.check(L.lib.phx_rollout_host(env._handle, T, a.data_ptr(), None, h_obs.data_ptr(),
                                           None, h_rew.data_ptr(), None, None, None,
                                           h_all.data_ptr()))
            assert torch.equal(h_obs, o), chunks
            assert torch.equal(h_rew, rw), chunks
            assert torch.equal(h_all, ad), chunks
        assert np.array_equal(shop(dev), shop(env)), chunks
        env.check_errors()
        env.close()
    dev.check_errors()
    dev.close()
