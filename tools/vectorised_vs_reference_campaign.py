#!/usr/bin/env python
"""One-off campaign (build container, CPU): oracle/vectorised.py (the numpy restatement the bench
kernels are compared with at full size and at random parameters, tools/fuzz_campaign5.py) against
the supply-chain workload executed by the UNMODIFIED reference, at random parameter sets --
customers 1-6, max order 2-9, max stock 10-1000, episode length 3-40, two episodes, negative /
oversized / missing actions.  Needs /root/reference.

    python tools/vectorised_vs_reference_campaign.py [--count 300]
"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=300)
    a = ap.parse_args()
    from oracle import harness, ref_shim, vectorised
    from oracle.workloads import supply_chain as wl

    ref = ref_shim.import_reference()
    bad = []
    for c in range(a.count):
        r = np.random.RandomState(130000 + c)
        nc = int(r.choice([5, 5, 1, 2, 3, 4, 6]))
        max_order = int(r.choice([5, 5, 2, 3, 4, 6, 7, 9]))
        max_stock = int(r.choice([100, 100, 10, 37, 250, 1000]))
        T, n_ep, n_env = int(r.randint(3, 41)), 2, 3
        seed, off = int(r.randint(1, 1 << 30)), int(r.choice([0, 7, 1 << 20]))
        A = r.uniform(-20, 1.3 * max_stock, size=(n_env, n_ep, T, 1)).astype(np.float32)
        M = (r.uniform(size=(n_env, n_ep, T)) > 0.15).astype(np.uint8)
        v = vectorised.SupplyChainVec(n_env, seed, n_customers=nc, max_order=max_order,
                                      max_stock=max_stock, num_steps=T, env_offset=off)
        vec = {"reset_obs": [], "obs": [], "reward": [], "all_trunc": [], "state": []}
        for ep in range(n_ep):
            vec["reset_obs"].append(v.reset())
            for t in range(T):
                o = v.step(A[:, ep, t, 0], M[:, ep, t])
                for k in ("obs", "reward", "all_trunc", "state"):
                    vec[k].append(np.array(o[k]))
        ok = True
        for e in range(n_env):
            st = wl.order_stream(seed, off + e, n_customers=nc, max_order=max_order)
            env = wl.build(ref, st, n_customers=nc, max_order=max_order, max_stock=max_stock, num_steps=T)
            tr = harness.run_supply_chain(env, harness.EpisodeClock([st]), A[e], M[e])
            for ep in range(n_ep):
                ok &= np.array_equal(tr["reset_obs"][ep], vec["reset_obs"][ep][e])
                for t in range(T):
                    i = ep * T + t
                    ok &= np.array_equal(tr["obs"][ep, t], vec["obs"][i][e])
                    ok &= tr["reward"][ep, t] == vec["reward"][i][e]
                    ok &= tr["all_trunc"][ep, t] == vec["all_trunc"][i][e]
                    ok &= np.array_equal(tr["state"][ep, t], vec["state"][i][e])
        if not ok:
            bad.append((c, nc, max_order, max_stock, T))
            print("MISMATCH", bad[-1], flush=True)
    print(json.dumps({"parameter_sets": a.count, "episodes": a.count * 6, "mismatches": bad}))


if __name__ == "__main__":
    main()
