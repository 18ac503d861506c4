"""Run-time specialisation (phx_jit_source / phx_load_specialised, phantom_b200/jit.py): the
step kernel rebuilt with one handle's env class as a compile-time constant must reproduce the
generic kernel -- and therefore the reference goldens -- bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from .generic_parity import assert_batchsteps_equal, run_device_vs_golden  # noqa: E402


def test_specialised_stackelberg_matches_reference_golden(golden_dir):
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    g = np.load(os.path.join(golden_dir, "stackelberg_reference.npz"))

    def make(**kw):
        env = StackelbergGameEnv(exec_mode="thread", **kw)
        env.specialise()
        assert "specialised" in env.exec_name
        return env

    run_device_vs_golden(make, g).close()


def test_specialised_simple_market_matches_reference_golden(golden_dir):
    """FSM + env-level words + float64 state through the specialised build."""
    from phantom_b200.envs import simple_market as sm

    g = np.load(os.path.join(golden_dir, "simple_market_wide_reference.npz"))
    buyers, n_sellers, T = [tuple(b) for b in g["buyers"]], int(g["n_sellers"]), g["actions"].shape[2]

    def make(**kw):
        env = sm.example_env(buyers, n_sellers, T, exec_mode="thread", **kw)
        return env.specialise()

    from .test_gpu_simple_market import market_state

    run_device_vs_golden(make, g, state_fn=market_state).close()


def test_specialised_tile_engine_matches_reference_golden(golden_dir):
    """The lane-per-agent tiling (C3, 32 agents, three FSM stages) specialised: flags, stage tables
    and masks fold; the golden of the unmodified reference still matches bit for bit."""
    from phantom_b200.envs.market import MarketEnv

    g = np.load(os.path.join(golden_dir, "market_reference.npz"))

    def make(**kw):
        env = MarketEnv(**kw).specialise()
        assert env.exec_name == "queue(G=32, collective, specialised)"
        return env

    run_device_vs_golden(make, g).close()


def test_specialised_equals_generic_at_scale_and_caches():
    """Auto-reset rollouts at 16 384 envs: specialised == generic; the second specialise() of an
    identical handle is a cache hit (no compile); an env class without a specialisation says so."""
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv
    from phantom_b200.envs.supply_chain import SupplyChainEnv
    from phantom_b200.envs.supply_chain2 import SupplyChain2Env

    for make, S in ((lambda: StackelbergGameEnv(num_envs=16384, seed=3, auto_reset=True, exec_mode="thread"), 4),
                    (lambda: SupplyChain2Env(num_envs=16384, seed=3, auto_reset=True), 2)):
        A = np.random.RandomState(1).uniform(0, 1, size=(60, 16384, S, 1)).astype(np.float32)
        if S == 2:
            A *= 100
        a, b = make(), make()
        a.reset_batch(); b.reset_batch()
        b.specialise()
        assert_batchsteps_equal(a.rollout_batch(A), b.rollout_batch(A))
        c = make()
        c.reset_batch()
        from phantom_b200 import jit

        before = sorted(os.listdir(jit.CACHE))
        c.specialise()
        assert sorted(os.listdir(jit.CACHE)) == before, "identical handle: the cubin must come from the cache"
        for e in (a, b, c):
            e.check_errors()
            e.close()
    fast = SupplyChainEnv(num_envs=64)
    fast.reset_batch()
    with pytest.raises(RuntimeError, match="no run-time specialisation"):
        fast.specialise()  # the schedule-specialised fast kernel has nothing left to fold
    fast.close()


def test_static_schedule_supply_chain_and_stackelberg(monkeypatch):
    """Specialised units of the thread-per-env engine carry a STATIC message schedule when the
    device program declares its sends (csrc/phx_engine.cuh StaticPlan): every potential message
    has a fixed slot and the handlers run in the planned order.  Must equal the dynamic queue --
    with actions missing (conditional sends -> invalid slots), auto-reset wraps, and on a
    supply chain whose agent order differs from the example's."""
    import phantom_b200 as ph
    from phantom_b200.envs import supply_chain as sc
    from phantom_b200.envs.stackelberg_game import StackelbergGameEnv

    def shuffled_chain(**kw):  # customers first, then the shop, the factory last
        ids = [f"CUST{i + 1}" for i in range(3)]
        agents = [sc.CustomerAgent(c, "SHOP") for c in ids] + [sc.ShopAgent("SHOP", "WAREHOUSE"),
                                                               sc.FactoryAgent("WAREHOUSE")]
        net = ph.Network(agents)
        net.add_connection("SHOP", "WAREHOUSE")
        net.add_connections_between(["SHOP"], ids)
        return ph.PhantomEnv(num_steps=9, network=net, exec_mode="thread", **kw)

    cases = ((lambda **kw: StackelbergGameEnv(exec_mode="thread", num_steps=10, **kw), 4, 1.0),
             (lambda **kw: sc.SupplyChainEnv(exec_mode="thread", num_steps=9, **kw), 1, 100.0),
             (shuffled_chain, 1, 100.0))
    for make, S, scale in cases:
        E, T = 4096, 47
        r = np.random.RandomState(S)
        A = (r.uniform(0, 1, size=(T, E, S, 1)) * scale).astype(np.float32)
        M = (r.uniform(size=(T, E, S)) > 0.15).astype(np.uint8)
        envs = []
        for static in ("1", "0", None):
            env = make(num_envs=E, seed=5, auto_reset=True)
            env.reset_batch()
            if static is not None:
                monkeypatch.setenv("PHX_JIT_STATIC_PLAN", static)
                env.specialise()
                assert ("static schedule" in env.exec_name) == (static == "1"), env.exec_name
            envs.append(env)
        outs = [[x.clone() for x in env.rollout_batch(A, M)] for env in envs]
        from phantom_b200 import BatchStep

        for o in outs[:2]:
            assert_batchsteps_equal(BatchStep(*o), BatchStep(*outs[2]))
        outs = [[x.clone() for x in env.rollout_batch(A)] for env in envs]  # every action present
        for o in outs[:2]:
            assert_batchsteps_equal(BatchStep(*o), BatchStep(*outs[2]))
        for env in envs:
            env.check_errors()
            env.close()
