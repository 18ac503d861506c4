#!/usr/bin/env python
"""One-off campaign (build container, CPU): the oracle-side runs that tools/fuzz_campaign2.py and
fuzz_campaign4.py compare the device with -- the oracle port's restatement of simple_market, and
the authored reference-API workloads (supply_chain2, market, stackelberg, dense) on the oracle
port -- against the UNMODIFIED reference (for simple_market: the unmodified example classes) on
the SAME random configurations.  Closes the chain device == oracle == reference for those
campaigns.  Needs /root/reference.

    python tools/ports_vs_reference_campaign.py [--count 60]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def same(a, b):
    keys = [k for k in a if k != "messages"]
    return all(np.array_equal(np.asarray(a[k]), np.asarray(b[k])) for k in keys)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=60)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle import harness, ref_shim, rng
    from oracle.make_golden import generic_actions
    from oracle.workloads import dense as wd
    from oracle.workloads import market as wm
    from oracle.workloads import simple_market as wsm
    from oracle.workloads import stackelberg as ws
    from oracle.workloads import supply_chain2 as w2

    ref = ref_shim.import_reference()
    bad, runs = [], 0

    def both(tag, cfg, run):
        nonlocal runs
        runs += 1
        if not same(run(po), run(ref)):
            bad.append((tag, cfg))
            print("MISMATCH", tag, cfg, flush=True)

    for c in range(a.count):
        # ---- simple_market (campaign 2's generator): port restatement vs the unmodified example
        r = np.random.RandomState(9000 + c)
        n_sellers = int(r.randint(1, 16))
        n_buyers = int(r.randint(1, 31 if r.uniform() < 0.5 else 6))
        buyers = tuple((float(np.round(p, 3)), float(np.round(lo, 3)), float(np.round(lo + w, 3)))
                       for p, lo, w in zip(r.uniform(0.1, 0.95, n_buyers), r.uniform(0, 0.6, n_buyers),
                                           r.uniform(0, 0.4, n_buyers)))
        T, seed = int(r.randint(4, 11)), 40000 + c
        A, M = wsm.actions_for(1, 2, T, n_buyers, n_sellers, c)

        def run_sm(ph):
            coords = wsm.Coords(seed, 0)
            with wsm.contract_rng(coords, {f"b{i + 1}": i for i in range(n_buyers)}):
                if ph is po:
                    env, _ = wsm.build(po, po.utils.samplers.UniformFloatSampler, buyers, n_sellers, T)
                else:
                    env, _ = wsm.build_reference(buyers, n_sellers, T)
                return harness.run_generic(env, harness.EpisodeClock([coords]), A[0], M[0], wsm.OBS_DIM,
                                           state_fn=wsm.state, convert=wsm.to_action(env))
        both("simple_market", (n_buyers, n_sellers, T), run_sm)

        # ---- supply_chain2 on a StochasticNetwork (campaign 2b's generator)
        r = np.random.RandomState(11000 + c)
        n_shops = int(r.randint(1, 4))
        n_cust = int(r.randint(1, 8 - n_shops))
        rates = (float(np.round(r.uniform(0.3, 1.0), 3)), float(np.round(r.uniform(0.2, 1.0), 3)))
        shuffle = bool(r.uniform() < 0.5)
        T, seed = int(r.randint(5, 21)), 50000 + c
        A2 = r.uniform(0, 100, size=(4, 2, T, n_shops, 1)).astype(np.float32)
        M2 = (r.uniform(size=(4, 2, T, n_shops)) > 0.1).astype(np.uint8)

        def run_c2(ph):
            streams = {s: rng.StepStream(seed, 0, s)
                       for s in (w2.STREAM_ORDER, w2.STREAM_SAMPLER, w2.STREAM_SHOP_CHOICE, w2.STREAM_CONNECTIVITY)}
            with harness.patched_np_uniform(streams[w2.STREAM_SAMPLER]), \
                    harness.patched_np_random(streams[w2.STREAM_CONNECTIVITY]):
                env = w2.build(ph, streams, ph.utils.samplers.UniformFloatSampler, n_shops=n_shops,
                               n_customers=n_cust, num_steps=T, rates=rates, shuffle_batches=shuffle)
                clock = harness.EpisodeClock(list(streams.values()))
                slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
                with harness.patched_np_shuffle(seed, 0, clock, env, slot_of):
                    return harness.run_generic(env, clock, A2[0], M2[0], 4, state_fn=w2.state)
        both("supply_chain2", (n_shops, n_cust, rates, shuffle, T), run_c2)

        # ---- C3 / C4 / C5 workloads (campaign 4's generator)
        r = np.random.RandomState(60000 + c)
        nm, nt = int(r.randint(1, 8)), int(r.randint(1, 25))
        T, seed, shuffle = int(r.randint(4, 25)), 70000 + c, bool(r.uniform() < 0.3)
        Am, Mm = generic_actions(3, 2, T, nm + nt, discrete_from=nm, seed=c, p_missing=0.1)

        def run_m(ph):
            st = rng.StepStream(seed, 0, wm.STREAM_TAKER_VALUE)
            env = wm.build(ph, st, n_makers=nm, n_takers=nt, num_steps=T, shuffle_batches=shuffle)
            clock = harness.EpisodeClock([st])
            slot_of = {aid: i for i, aid in enumerate(env.agent_ids)}
            with harness.patched_np_shuffle(seed, 0, clock, env, slot_of):
                return harness.run_generic(env, clock, Am[0], Mm[0], 3)
        both("market", (nm, nt, T, shuffle), run_m)
        nf = int(r.randint(1, 8))
        T, seed = int(r.randint(4, 31)), 80000 + c
        As, Ms = generic_actions(3, 2, T, nf + 1, seed=c + 1, p_missing=0.1)

        def run_s(ph):
            st = rng.StepStream(seed, 0, ws.STREAM_FOLLOWER_VALUE)
            env = ws.build(ph, st, n_followers=nf, num_steps=T)
            return harness.run_generic(env, harness.EpisodeClock([st]), As[0], Ms[0], 2, state_fn=ws.state)
        both("stackelberg", (nf, T), run_s)
        n = int(r.choice([2, 3, 5, 8, 12, 17, 31, 32, 33, 64, 100, 127, 128]))
        dens = float(r.choice([1.0, 0.5, 0.1]))
        up = np.triu((r.uniform(size=(n, n)) < dens).astype(np.int64), 1)
        adj = up + up.T
        T = int(r.randint(2, 7))
        rl = [2, 2, 3, None][int(r.randint(4))]
        Ad, Md = generic_actions(3, 2, T, n, seed=c + 2, p_missing=0.1 if r.uniform() < 0.5 else 0.0)

        def run_d(ph):
            env = wd.build(ph, n_agents=n, adjacency=adj, num_steps=T, round_limit=rl)
            return harness.run_generic(env, harness.EpisodeClock([]), Ad[0], Md[0], 3, state_fn=wd.state)
        both("dense", (n, dens, T, rl), run_d)
    print(json.dumps({"configurations": runs, "mismatches": bad}))


if __name__ == "__main__":
    main()
