"""SURVEY 8(f) rows 2 and 4 on the device: StochasticNetwork (per-env edges resampled at every
reset, phantom/network.py:340-453) and BatchResolver(shuffle_batches=True)
(phantom/resolvers.py:150-151), against fixtures produced by the UNMODIFIED reference under the
RNG contract (stream 5 connectivity, streams 0x100 + receiver for the shuffles)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from .generic_parity import (assert_batchsteps_equal, run_device_trace_vs_golden,  # noqa: E402
                             run_device_vs_golden)
from .test_gpu_market import market_state  # noqa: E402
from .test_gpu_supply_chain2 import shop_state  # noqa: E402


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_market_with_shuffled_batches_matches_reference(golden_dir):
    from phantom_b200.envs.market import MarketEnv

    g = load(golden_dir, "market_shuffle_reference.npz")
    make = lambda **kw: MarketEnv(shuffle_batches=True, **kw)
    run_device_vs_golden(make, g, market_state, state_every=4).close()
    run_device_trace_vs_golden(make, g, 1)
    # and the shuffle is not a no-op on this workload
    plain = load(golden_dir, "market_reference.npz")
    assert not np.array_equal(g["state"][0, 0, :, :, :], plain["state"][0, 0, :, :, :])


@pytest.mark.parametrize("variant,exec_mode", [("shuffle", "queue"), ("plain", "queue"),
                                               ("plain", "thread"), ("shuffle", "wide"),
                                               ("plain", "wide")])
def test_stochastic_network_matches_reference(golden_dir, variant, exec_mode):
    """Every episode's graph, every step's outputs and state, and the message trace (which shows
    both the dropped deliveries and the shuffled batch order)."""
    from phantom_b200.envs.supply_chain2 import SupplyChain2Env

    shuffle = variant == "shuffle"
    g = load(golden_dir, "supply_chain2_stochastic_reference.npz" if shuffle
             else "supply_chain2_stochastic_plain_reference.npz")
    rates = tuple(g["rates"])
    T = g["actions"].shape[2]
    make = lambda **kw: SupplyChain2Env(num_steps=T, rates=rates, shuffle_batches=shuffle,
                                        exec_mode=exec_mode, **kw)
    seed, A, M = int(g["seed"]), g["actions"], g["action_mask"]
    env = make(num_envs=A.shape[0], seed=seed)
    assert ("thread" in env.exec_name) == (exec_mode == "thread")
    assert ("wide" in env.exec_name) == (exec_mode == "wide")  # the 128-lane block engine, forced
    from .generic_parity import assert_device_step_equal

    for ep in range(A.shape[1]):
        obs, mask = env.reset_batch()
        assert np.array_equal(env.adjacency(), g["adjacency"][:, ep]), f"graphs of episode {ep}"
        assert np.array_equal(obs.cpu().numpy(), g["reset_obs"][:, ep])
        for t in range(T):
            out = env.step_batch(A[:, ep, t], M[:, ep, t])
            assert_device_step_equal(out, g, ep, t, f"ep {ep} t {t}")
            if t % 5 == 0 or t == T - 1:
                assert np.array_equal(shop_state(env), g["state"][:, ep, t]), (ep, t)
    env.check_errors()
    env.close()
    run_device_trace_vs_golden(make, g, 4)


@pytest.mark.parametrize("exec_mode", ["thread", "queue", "wide"])
def test_stochastic_network_auto_reset_and_sharding(exec_mode):
    """Auto-reset inside a rollout resamples the graphs exactly like explicit resets; results do
    not depend on how envs are split over handles; edge frequencies follow the rates."""
    import torch

    from phantom_b200 import BatchStep
    from phantom_b200.envs.supply_chain2 import SupplyChain2Env

    E, T, steps, seed, rates = 4096, 10, 35, 77, (0.5, 0.25)
    mk = lambda n, off=0, **kw: SupplyChain2Env(num_envs=n, seed=seed, env_offset=off, num_steps=T,
                                                rates=rates, exec_mode=exec_mode, **kw)
    A = np.random.RandomState(5).uniform(0, 100, size=(steps, E, 2, 1)).astype(np.float32)
    auto, manual = mk(E, auto_reset=True), mk(E)
    lo, hi = mk(E // 2, auto_reset=True), mk(E // 2, E // 2, auto_reset=True)
    for e in (auto, manual, lo, hi):
        e.reset_batch()
    ro = auto.rollout_batch(A)
    parts = [lo.rollout_batch(A[:, : E // 2]), hi.rollout_batch(A[:, E // 2:])]
    assert_batchsteps_equal(ro, BatchStep(*[torch.cat([y, z], dim=1) for y, z in zip(*parts)]))
    for t in range(steps):
        out = manual.step_batch(A[t])
        assert np.array_equal(ro.rewards[t].cpu().numpy(), out.rewards.cpu().numpy()), t
        if out.all_done.cpu().numpy()[0, 1]:
            obs, _ = manual.reset_batch()
            assert np.array_equal(ro.observations[t].cpu().numpy(), obs.cpu().numpy()), t
        else:
            assert np.array_equal(ro.observations[t].cpu().numpy(), out.observations.cpu().numpy()), t
    adj = auto.adjacency()
    assert np.array_equal(adj, manual.adjacency())
    assert np.array_equal(adj, np.concatenate([lo.adjacency(), hi.adjacency()]))
    assert np.array_equal(adj, adj.transpose(0, 2, 1))
    # WAREHOUSE = slot 0, shops 1-2, customers 3-7
    assert abs(adj[:, 1:3, 0].mean() - rates[0]) < 0.02
    assert abs(adj[:, 1:3, 3:].mean() - rates[1]) < 0.02
    assert adj[:, 3:, 3:].sum() == 0 and adj[:, 0, 3:].sum() == 0
    for e in (auto, manual, lo, hi):
        e.check_errors()
        e.close()


def test_unsupported_combinations_fail_loudly():
    import phantom_b200 as ph
    from phantom_b200 import _lib as L
    from phantom_b200.envs.dense import DenseEnv
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    # the schedule-specialised kernel assumes a static graph and push-order batches
    env = SupplyChainEnv(num_envs=4, exec_mode="fast")
    env.network.resolver.shuffle_batches = True
    with pytest.raises(L.PhxError, match="shuffle"):
        env.reset_batch()
    # ... while auto mode falls back to the generic engine
    env = SupplyChainEnv(num_envs=4)
    env.network.resolver.shuffle_batches = True
    env.reset_batch()
    assert "queue" in env.exec_name
    env.close()
    env = DenseEnv(n_agents=12, num_envs=2)
    env.network.resolver.shuffle_batches = True
    with pytest.raises(L.PhxError, match="queue engine"):
        env.reset_batch()
