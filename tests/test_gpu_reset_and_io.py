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


@pytest.mark.parametrize("lo,hi,expect_wire", [(0.0, 100.0, True), (-40.0, 140.0, True),
                                               (-60000.0, 100.0, False)])
def test_rollout_host_compact_wire_equals_device_planes(lo, hi, expect_wire, monkeypatch):
    """The lean host-buffer call of the supply chain moves ONE 32-bit word per env-step across
    PCIe and expands it on host threads (csrc/phx_sc_wire.*): the float32 planes must equal the
    device path's bit for bit -- in distribution, with negative stock / sales (values outside the
    wire fields: the call falls back to the float planes the kernel also wrote), across
    auto-reset wraps and over two consecutive calls."""
    import torch

    from phantom_b200 import _lib as L
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    E, T, seed = 8192, 29, 14
    r = np.random.RandomState(7)
    tapes = [r.uniform(lo, hi, size=(T, E, 1, 1)).astype(np.float32) for _ in range(2)]
    dev = SupplyChainEnv(num_envs=E, seed=seed, num_steps=13, auto_reset=True)
    host = SupplyChainEnv(num_envs=E, seed=seed, num_steps=13, auto_reset=True)
    monkeypatch.setenv("PHX_NO_WIRE", "1")
    plain = SupplyChainEnv(num_envs=E, seed=seed, num_steps=13, auto_reset=True)
    for env in (dev, host, plain):
        env.reset_batch()
    h_obs, h_rew, h_all = _host_planes(T, E)
    p_obs, p_rew, p_all = _host_planes(T, E)
    for A in tapes:
        want = dev.rollout_batch(A)
        a = torch.as_tensor(A).pin_memory()
        for env, (o, rw, ad) in ((host, (h_obs, h_rew, h_all)), (plain, (p_obs, p_rew, p_all))):
            o.fill_(-1.0); rw.fill_(-1.0); ad.fill_(7)
            L.check(L.lib.phx_rollout_host(env._handle, T, a.data_ptr(), None, o.data_ptr(), None,
                                           rw.data_ptr(), None, None, None, ad.data_ptr()))
            assert torch.equal(o, want.observations.cpu())
            assert torch.equal(rw, want.rewards.cpu())
            assert torch.equal(ad, want.all_done.cpu())
    if not expect_wire:  # the out-of-range tape really left the wire fields
        stock = np.asarray(dev.agents["SHOP"].stock)
        assert stock.min() < -32768 or np.asarray(dev.agents["SHOP"].sales).min() < 0
    shop = lambda env: np.stack([np.asarray(getattr(env.agents["SHOP"], c))
                                 for c in ("stock", "sales", "missed_sales")], 1)
    assert np.array_equal(shop(dev), shop(host)) and np.array_equal(shop(dev), shop(plain))
    for env in (dev, host, plain):
        env.check_errors()
        env.close()
