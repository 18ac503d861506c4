"""Differential fuzz on the B200: random supply-chain topologies (agent order, customer count,
dropped edges with ignore_connection_errors, missing actions) stepped by BOTH variants of the
generic engine and compared, message by message, with the object-level oracle running the same
env definition.  Seeds are fixed, so failures reproduce."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.phantom_oracle as po  # noqa: E402
from oracle import harness, rng  # noqa: E402
from oracle.workloads import supply_chain as wl  # noqa: E402


def build_pair(case_seed: int, exec_mode: str):
    import phantom_b200 as ph
    from phantom_b200.envs import supply_chain as sc

    r = np.random.RandomState(case_seed)
    nc = int(r.randint(1, 7))
    ids = ["SHOP", "WAREHOUSE"] + [f"CUST{i + 1}" for i in range(nc)]
    order = list(r.permutation(ids))
    drop = [c for c in ids[2:] if r.uniform() < 0.25]  # customers whose edge to the shop is cut
    drop_factory = r.uniform() < 0.15
    seed = int(r.randint(1, 1 << 30))

    def make_dev():
        agents = {"SHOP": sc.ShopAgent("SHOP", "WAREHOUSE"), "WAREHOUSE": sc.FactoryAgent("WAREHOUSE")}
        agents.update({c: sc.CustomerAgent(c, "SHOP") for c in ids[2:]})
        net = ph.Network([agents[a] for a in order], ph.resolvers.BatchResolver(enable_tracking=True),
                         ignore_connection_errors=True)
        if not drop_factory:
            net.add_connection("SHOP", "WAREHOUSE")
        net.add_connections_between(["SHOP"], [c for c in ids[2:] if c not in drop])
        env = ph.PhantomEnv(num_steps=12, network=net, seed=seed, exec_mode=exec_mode)
        env.max_order, env.max_stock = 5, 100
        return env

    st = wl.order_stream(seed, 0, n_customers=nc)
    ref = wl.build(po, st, n_customers=nc, num_steps=12, enable_tracking=True)
    ref.network.ignore_connection_errors = True
    ref.network.agents = {k: ref.network.agents[k] for k in order}
    g = ref.network.graph
    for c in drop:
        del g._succ["SHOP"][c], g._succ[c]["SHOP"]
    if drop_factory:
        del g._succ["SHOP"]["WAREHOUSE"], g._succ["WAREHOUSE"]["SHOP"]
    return make_dev(), ref, harness.EpisodeClock([st]), r, order


def msgs(env):
    return [(m.sender_id, m.receiver_id, type(m.payload).__name__, m.payload.size)
            for m in env.network.resolver.tracked_messages]


@pytest.mark.parametrize("exec_mode", ["thread", "queue"])
@pytest.mark.parametrize("case_seed", range(12))
def test_random_supply_chain_topology(case_seed, exec_mode):
    env, ref, clock, r, order = build_pair(case_seed, exec_mode)
    for ep in range(2):
        clock.on_reset()
        o_ref, _ = ref.reset()
        o, _ = env.reset()
        assert np.array_equal(o["SHOP"], o_ref["SHOP"])
        for t in range(12):
            a = {} if r.uniform() < 0.2 else {"SHOP": r.uniform(-5, 140, size=(1,)).astype(np.float32)}
            clock.on_step(ref)
            ref.network.resolver.clear_tracked_messages()
            env.network.resolver.clear_tracked_messages()
            s_ref, s = ref.step(a), env.step(a)
            assert np.array_equal(s.observations["SHOP"], s_ref.observations["SHOP"]), (order, ep, t)
            assert s.rewards["SHOP"] == np.float32(s_ref.rewards["SHOP"]), (order, ep, t)
            assert s.truncations == s_ref.truncations
            assert msgs(env) == msgs(ref), (order, ep, t)
    env.close()
