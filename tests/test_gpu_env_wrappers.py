"""SingleAgentEnvAdapter (reference: phantom/env_wrappers.py:23-196) over device envs: the
adapter's view of one agent == driving the same env by hand with the same policies."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _policies():
    import phantom_b200 as ph

    class BuyerPolicy(ph.Policy):  # examples/environments/simple_market/base_policy.py:4-11
        def compute_action(self, obs):
            obs = np.asarray(obs)
            return np.where((obs[..., 1] != 0) & (obs[..., 0] <= obs[..., 2]), obs[..., 1], 0.0)

    class FixedPrice(ph.Policy):
        def __init__(self, observation_space, action_space, price):
            super().__init__(observation_space, action_space)
            self.price = price

        def compute_action(self, obs):
            return np.full(np.asarray(obs).shape[:-1], self.price, np.float32)

    return BuyerPolicy, FixedPrice


def test_adapter_single_env_matches_manual_loop():
    import phantom_b200 as ph
    from phantom_b200.envs import simple_market as sm

    BuyerPolicy, FixedPrice = _policies()
    others = {"b1": (BuyerPolicy, {}), "b2": (BuyerPolicy, {}), "b3": (BuyerPolicy, {}),
              "s2": (FixedPrice, {"price": 0.375})}
    ad = ph.SingleAgentEnvAdapter(sm.example_env, "s1", others, {"seed": 3, "num_steps": 10})
    assert ad.active_agent == "s1" and ad.n_agents == 5 and ad.action_space.shape == (1,)
    twin = sm.example_env(seed=3, num_steps=10)
    pol = {aid: cls(twin[aid].observation_space, twin[aid].action_space, **cfg)
           for aid, (cls, cfg) in others.items()}
    obs, _ = twin.reset()
    o0, _ = ad.reset()
    assert np.array_equal(o0, obs["s1"])
    prices = np.random.RandomState(0).uniform(0, 1, 10).astype(np.float32)
    for t in range(10):
        acts = {aid: p.compute_action(obs[aid]) for aid, p in pol.items() if aid in obs}
        acts["s1"] = prices[t]
        step = twin.step(acts)
        obs = step.observations
        o, r, term, trunc, info = ad.step(prices[t])
        assert (o is None) == ("s1" not in step.observations)
        if o is not None:
            assert np.array_equal(o, step.observations["s1"]) and r == step.rewards["s1"]
        assert term == step.terminations.get("s1") and trunc == step.truncations.get("s1")
    ad.close()
    twin.close()


def test_adapter_supply_chain_single_agent():
    import phantom_b200 as ph
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    ad = ph.SingleAgentEnvAdapter(SupplyChainEnv, "SHOP", {}, {"seed": 5})
    twin = SupplyChainEnv(seed=5)
    twin.reset()
    for a in (10.0, 55.5, 3.0):
        step = twin.step({"SHOP": [a]})
        o, r, term, trunc, _ = ad.step([a])
        assert np.array_equal(o, step.observations["SHOP"]) and r == step.rewards["SHOP"]
        assert term is False and trunc is False
    ad.close()
    twin.close()


def test_adapter_vector_env():
    """env_config num_envs = E: a vector env for the selected agent, policies see [E, obs] batches."""
    import phantom_b200 as ph
    from phantom_b200.envs import simple_market as sm

    BuyerPolicy, FixedPrice = _policies()
    E = 64
    others = {"b1": (BuyerPolicy, {}), "b2": (BuyerPolicy, {}), "b3": (BuyerPolicy, {}),
              "s2": (FixedPrice, {"price": 0.5})}
    ad = ph.SingleAgentEnvAdapter(sm.example_env, "s1", others,
                                  {"seed": 3, "num_steps": 10, "num_envs": E})
    singles = [ph.SingleAgentEnvAdapter(sm.example_env, "s1", others,
                                        {"seed": 3, "num_steps": 10, "env_offset": e})
               for e in (0, 17, 63)]
    prices = np.random.RandomState(1).uniform(0, 1, (10, E)).astype(np.float32)
    for t in range(10):
        o, r, term, trunc, _ = ad.step(prices[t])
        assert o.shape == (E, 2) and r.shape == (E,)
        for k, e in enumerate((0, 17, 63)):
            so, sr, st, su, _ = singles[k].step(prices[t, e])
            if so is not None:
                assert np.array_equal(so, o[e]) and sr == r[e]
    ad.close()
    for s in singles:
        s.close()
