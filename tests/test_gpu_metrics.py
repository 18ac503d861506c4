"""Metrics on device envs: the reference's metric classes
(phantom/metrics.py) read agent attributes; on a device env those are state columns."""
from collections import defaultdict

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

import oracle.phantom_oracle as po  # noqa: E402
from oracle import harness, rng  # noqa: E402
from oracle.workloads import supply_chain as wl  # noqa: E402


def test_supply_chain_metrics_match_oracle_attributes():
    """The metrics dict of the reference example (supply_chain.py:178-182), episode-reduced."""
    import phantom_b200 as ph
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    metrics = {
        "SHOP/stock": ph.metrics.SimpleAgentMetric("SHOP", "stock", "mean"),
        "SHOP/sales": ph.metrics.SimpleAgentMetric("SHOP", "sales", "mean"),
        "SHOP/missed_sales": ph.metrics.SimpleAgentMetric("SHOP", "missed_sales", "sum", "last"),
        "step": ph.metrics.SimpleEnvMetric("current_step", "last"),
    }
    seed = 5
    env = SupplyChainEnv(seed=seed)
    st = wl.order_stream(seed, 0)
    ref = wl.build(po, st)
    clock = harness.EpisodeClock([st])
    clock.on_reset(); ref.reset(); env.reset()
    values, want = defaultdict(list), defaultdict(list)
    r = np.random.RandomState(0)
    for t in range(100):
        a = {"SHOP": r.uniform(0, 100, size=(1,)).astype(np.float32)}
        clock.on_step(ref); ref.step(a); env.step(a)
        ph.metrics.logging_helper(env, metrics, values)
        shop = ref.agents["SHOP"]
        want["SHOP/stock"].append(shop.stock)
        want["SHOP/sales"].append(shop.sales)
        want["SHOP/missed_sales"].append(shop.missed_sales)
        want["step"].append(ref.current_step)
    assert metrics["SHOP/stock"].reduce(values["SHOP/stock"], "train") == np.mean(want["SHOP/stock"])
    assert metrics["SHOP/sales"].reduce(values["SHOP/sales"], "train") == np.mean(want["SHOP/sales"])
    assert metrics["SHOP/missed_sales"].reduce(values["SHOP/missed_sales"], "train") == np.sum(want["SHOP/missed_sales"])
    assert metrics["SHOP/missed_sales"].reduce(values["SHOP/missed_sales"], "evaluate") == want["SHOP/missed_sales"][-1]
    assert metrics["step"].reduce(values["step"], "train") == 100
    assert np.array_equal(metrics["SHOP/stock"].reduce(values["SHOP/stock"], "evaluate"), want["SHOP/stock"])
    env.close()


@pytest.mark.parametrize("exec_mode", ["fast", "thread", "queue"])
def test_device_reduction_over_envs(exec_mode):
    """phx_reduce_field == numpy over the fetched column, for every kernel variant's layout."""
    import phantom_b200 as ph
    from phantom_b200.envs.supply_chain import SupplyChainEnv

    E = 50000
    env = SupplyChainEnv(num_envs=E, seed=3, exec_mode=exec_mode)
    env.reset_batch()
    A = np.random.RandomState(1).uniform(-30, 120, size=(7, E, 1, 1)).astype(np.float32)
    env.rollout_batch(A)
    m = ph.metrics.SimpleAgentMetric("SHOP", "stock")
    col = np.asarray(m.extract(env), np.int64)
    assert col.shape == (E,)
    assert m.extract_over_envs(env, "sum") == col.sum()
    assert m.extract_over_envs(env, "min") == col.min() and m.extract_over_envs(env, "max") == col.max()
    assert m.extract_over_envs(env, "mean") == pytest.approx(col.mean(), rel=1e-12)
    env.close()


def test_fsm_stage_filter_and_group_metric():
    """logging_helper's FSM stage filter (metrics.py:355-370) and AggregatedAgentMetric on the
    C3 market: total maker inventory is only recorded in the CLEARING stage."""
    import phantom_b200 as ph
    from phantom_b200.envs.market import MarketEnv

    env = MarketEnv(num_envs=64, seed=2)
    env.reset_batch()
    metrics = {
        "inventory": ph.metrics.AggregatedAgentMetric(env.maker_ids, "inventory", "sum", "last",
                                                      fsm_stages=["MAKER"]),
    }
    values = defaultdict(list)
    r = np.random.RandomState(0)
    for t in range(9):
        a = r.uniform(0, 1, size=(64, 31, 1)).astype(np.float32)
        a[:, 7:] = a[:, 7:] > 0.3
        env.step_batch(a)
        ph.metrics.logging_helper(env, metrics, values)
    rec = values["inventory"]
    # after step t the env is in stage (t+1) % 3; MAKER == 0 -> recorded after steps 2, 5, 8
    assert [v is not ph.metrics.not_recorded for v in rec] == [False, False, True] * 3
    last = metrics["inventory"].reduce(rec, "train")
    want = sum(np.asarray(env.agents[m].inventory) for m in env.maker_ids)
    assert last.shape == (64,) and np.array_equal(last, want) and (last <= 84).all()
    with pytest.raises(ValueError):
        ph.metrics.SimpleAgentMetric("a", "b", train_reduce_action="median")
    env.close()
