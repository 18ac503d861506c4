#!/usr/bin/env python
"""One-off campaign, part 1 (build container, CPU): random configurations of the reference's
digital_ads_market example -- advertiser counts per theme, auction strategy, episode length --
executed by the UNMODIFIED example file, written as fixtures to a scratch directory that travels
to the GPU box (tools/fuzz_campaign2.py compares the device against them).

    python tools/fuzz_campaign2_gen.py --out tests/_campaign --count 30
"""
import argparse
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "tests", "_campaign"))
    ap.add_argument("--count", type=int, default=30)
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--small-budgets", action="store_true",
                    help="budgets of 0.3 .. 2.5: most advertisers run out of money and terminate mid-episode")
    a = ap.parse_args()
    from oracle import harness
    from oracle.make_golden import pack_generic
    from oracle.workloads import digital_ads as wl

    os.makedirs(a.out, exist_ok=True)
    for c in range(a.first, a.first + a.count):
        r = np.random.RandomState(7000 + c)
        big = r.uniform() < 0.4
        counts = [int(r.randint(1, 41 if big else 11)) for _ in range(3)]
        strategy = "first" if r.uniform() < 0.5 else "second"
        T = int(r.randint(8, 21) if a.small_budgets else r.randint(4, 15))
        n_env, n_ep, seed = 2, 2, 30000 + c
        theme = {"travel": counts[0], "tech": counts[1], "sport": counts[2]}
        budgets = []
        for n in counts:
            lo = float(np.round(r.uniform(0.3, 1.5) if a.small_budgets else r.uniform(1.0, 12.0), 2))
            hi = float(np.round(lo + (r.uniform(0.1, 1.0) if a.small_budgets else r.uniform(0.5, 10.0)), 2))
            budgets += [(lo, hi, lo, float(np.round(hi - r.uniform(0.0, 0.4), 2)))] * n
        S = sum(counts)
        actions, mask = wl.actions_for(n_env, n_ep, T, S, seed % 1000)
        per_env = []
        for e in range(n_env):
            coords = wl.Coords(seed, e)
            with wl.contract_rng(coords):
                env = wl.build_reference(theme, budgets, T, strategy)
                tr = harness.run_generic(env, harness.EpisodeClock([coords]), actions[e], mask[e],
                                         wl.OBS_DIM, state_fn=wl.state, flatten=wl.flatten_obs)
            tr["messages"] = []
            per_env.append(tr)
        out = pack_generic(per_env, actions, mask, seed, 0, {})
        out["counts"], out["budgets"] = np.array(counts, np.int64), np.array(budgets, np.float64)
        out["second_price"] = np.int64(strategy == "second")
        np.savez_compressed(os.path.join(a.out, f"ads_{c:03d}.npz"), **out)
        print(c, counts, strategy, T, "terminated", int((out["term"] == 1).sum()), flush=True)


if __name__ == "__main__":
    main()
