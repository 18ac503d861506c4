#!/usr/bin/env python
"""One-off campaign, part 2 (GPU box): (a) the device against the digital_ads_market fixtures of
tools/fuzz_campaign2_gen.py on every tiling that fits; (b) random simple_market casts (1-30
buyers, 1-15 sellers) on the device against the oracle port's restatement of the example.

    python tools/fuzz_campaign2.py --dir tests/_campaign --market 60
"""
import argparse
import glob
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dir", default=os.path.join(REPO, "tests", "_campaign"))
    ap.add_argument("--market", type=int, default=60)
    ap.add_argument("--chain2", type=int, default=60)
    a = ap.parse_args()
    from tests.generic_parity import run_device_vs_golden
    from tests.test_gpu_digital_ads import ads_state
    from tests.test_gpu_simple_market import market_state
    from phantom_b200.envs import digital_ads_market as da
    from phantom_b200.envs import simple_market as sm
    from phantom_b200.utils.samplers import UniformFloatSampler

    report = {"ads": [], "market": [], "chain2": []}
    for path in sorted(glob.glob(os.path.join(a.dir, "ads_*.npz"))):
        g = np.load(path)
        counts = [int(x) for x in g["counts"]]
        n_agents = sum(counts) + 2
        modes = (["thread"] if n_agents <= 8 else []) + (["queue"] if n_agents <= 32 else []) + ["wide"]
        for mode in modes:
            def build(**kw):
                st = {f"ADV_{i + 1}": da.AdvertiserAgent.Supertype(budget=UniformFloatSampler(*b))
                      for i, b in enumerate(g["budgets"])}
                return da.DigitalAdsEnv(num_steps=g["actions"].shape[2],
                                        num_agents_theme={"travel": counts[0], "tech": counts[1], "sport": counts[2]},
                                        strategy="second" if int(g["second_price"]) else "first",
                                        agent_supertypes=st, exec_mode=mode, **kw)
            try:
                run_device_vs_golden(build, g, state_fn=ads_state).close()
                ok = True
            except Exception as exc:
                ok = False
                print("MISMATCH ads", os.path.basename(path), counts, mode, type(exc).__name__, str(exc)[:200], flush=True)
            report["ads"].append((os.path.basename(path), counts, mode, ok))
    import oracle.phantom_oracle as po
    from oracle import harness
    from oracle.workloads import simple_market as wl

    for c in range(a.market):
        r = np.random.RandomState(9000 + c)
        n_sellers = int(r.randint(1, 16))
        n_buyers = int(r.randint(1, 31 if r.uniform() < 0.5 else 6))
        buyers = tuple((float(np.round(p, 3)), float(np.round(lo, 3)), float(np.round(lo + w, 3)))
                       for p, lo, w in zip(r.uniform(0.1, 0.95, n_buyers), r.uniform(0, 0.6, n_buyers),
                                           r.uniform(0, 0.4, n_buyers)))
        T, n_env, n_ep, seed = int(r.randint(4, 11)), 3, 2, 40000 + c
        A, M = wl.actions_for(n_env, n_ep, T, n_buyers, n_sellers, c)
        per_env = []
        for e in range(n_env):
            coords = wl.Coords(seed, e)
            with wl.contract_rng(coords, {f"b{i + 1}": i for i in range(n_buyers)}):
                env, _ = wl.build(po, po.utils.samplers.UniformFloatSampler, buyers, n_sellers, T)
                per_env.append(harness.run_generic(env, harness.EpisodeClock([coords]), A[e], M[e],
                                                   wl.OBS_DIM, state_fn=wl.state, convert=wl.to_action(env)))
        g = {k: np.stack([t[k] for t in per_env]) for k in per_env[0]}
        g.update(actions=A, action_mask=M, seed=np.int64(seed))
        n_agents = n_buyers + n_sellers
        modes = (["thread"] if n_agents <= 8 else []) + (["queue"] if n_agents <= 32 else []) + ["wide"]
        for mode in modes:
            try:
                run_device_vs_golden(lambda **kw: sm.example_env(buyers, n_sellers, T, exec_mode=mode, **kw),
                                     g, state_fn=market_state).close()
                ok = True
            except Exception as exc:
                ok = False
                print("MISMATCH market", c, n_buyers, n_sellers, mode, type(exc).__name__, str(exc)[:200], flush=True)
            report["market"].append((c, n_buyers, n_sellers, mode, ok))
    # (c) the tutorial's multi-shop supply chain on a StochasticNetwork: random shop / customer
    # counts (<= 8 agents: the family's domain), connection rates, shuffled batches or not
    from oracle import rng
    from oracle.workloads import supply_chain2 as w2
    from phantom_b200.envs.supply_chain2 import SupplyChain2Env
    from tests.test_gpu_supply_chain2 import shop_state

    for c in range(a.chain2):
        r = np.random.RandomState(11000 + c)
        n_shops = int(r.randint(1, 4))
        n_cust = int(r.randint(1, 8 - n_shops))
        rates = (float(np.round(r.uniform(0.3, 1.0), 3)), float(np.round(r.uniform(0.2, 1.0), 3)))
        shuffle = bool(r.uniform() < 0.5)
        T, n_env, n_ep, seed = int(r.randint(5, 21)), 4, 2, 50000 + c
        A = r.uniform(0, 100, size=(n_env, n_ep, T, n_shops, 1)).astype(np.float32)
        M = (r.uniform(size=(n_env, n_ep, T, n_shops)) > 0.1).astype(np.uint8)
        per_env = []
        for e in range(n_env):
            streams = {s: rng.StepStream(seed, e, s)
                       for s in (w2.STREAM_ORDER, w2.STREAM_SAMPLER, w2.STREAM_SHOP_CHOICE, w2.STREAM_CONNECTIVITY)}
            with harness.patched_np_uniform(streams[w2.STREAM_SAMPLER]), \
                    harness.patched_np_random(streams[w2.STREAM_CONNECTIVITY]):
                env = w2.build(po, streams, po.utils.samplers.UniformFloatSampler, n_shops=n_shops,
                               n_customers=n_cust, num_steps=T, rates=rates, shuffle_batches=shuffle)
                clock = harness.EpisodeClock(list(streams.values()))
                slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
                with harness.patched_np_shuffle(seed, e, clock, env, slot_of):
                    per_env.append(harness.run_generic(env, clock, A[e], M[e], 4, state_fn=w2.state))
        g = {k: np.stack([t[k] for t in per_env]) for k in per_env[0] if k != "messages"}
        g.update(actions=A, action_mask=M, seed=np.int64(seed))
        for mode in (["queue", "wide"] if shuffle else ["thread", "queue", "wide"]):
            try:
                run_device_vs_golden(lambda **kw: SupplyChain2Env(n_shops, n_cust, num_steps=T, rates=rates,
                                                                  shuffle_batches=shuffle, exec_mode=mode, **kw),
                                     g, state_fn=shop_state).close()
                ok = True
            except Exception as exc:
                ok = False
                print("MISMATCH chain2", c, n_shops, n_cust, rates, shuffle, mode, type(exc).__name__, str(exc)[:200], flush=True)
            report["chain2"].append((c, n_shops, n_cust, mode, ok))
    bad = {k: [x for x in v if not x[-1]] for k, v in report.items()}
    print(json.dumps({"ads_runs": len(report["ads"]), "market_runs": len(report["market"]),
                      "chain2_runs": len(report["chain2"]), "mismatches": bad}))


if __name__ == "__main__":
    main()
