#!/usr/bin/env python
"""One-off campaign 5 (GPU box): the schedule-specialised supply-chain kernels (sc_fast_kernel /
sc_fast2_kernel, the bench kernels) at RANDOM parameters -- customers 1-6, max order 2-9, max
stock 10-1000, episode length 3-60, env counts 1-3000 (ragged around the 64-thread blocks),
launches of random length chained back to back (every clock phase, auto-reset wraps inside a
launch, programmatic dependent launch between them), negative / oversized / missing actions --
against oracle/vectorised.py (numpy, pinned to the reference fixtures): observations, rewards,
done flags of every step and the shop state after every launch; and the host-buffer entry point
(phx_rollout_host) against the device planes.

    python tools/fuzz_campaign5.py [--count 300]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--count", type=int, default=300)
    a = ap.parse_args()
    from oracle import vectorised
    from phantom_b200.envs import supply_chain as sc

    bad, kernels = [], {}
    for c in range(a.count):
        r = np.random.RandomState(120000 + c)
        nc = int(r.choice([5, 5, 5, 1, 2, 3, 4, 6]))
        max_order = int(r.choice([5, 5, 2, 3, 4, 6, 7, 9]))
        max_stock = int(r.choice([100, 100, 10, 37, 250, 1000]))
        num_steps = int(r.randint(3, 61))
        E = int(r.choice([1, 31, 33, 63, 64, 65, 96, 127, 128, 200, 512, 777, 1024, 3000]))
        seed, off = int(r.randint(1, 1 << 30)), int(r.choice([0, 0, 1 << 20]))
        masked = bool(r.uniform() < 0.3)
        env = sc.SupplyChainEnv(nc, num_steps=num_steps, num_envs=E, seed=seed, env_offset=off, auto_reset=True)
        env.max_order, env.max_stock = max_order, max_stock
        v = vectorised.SupplyChainVec(E, seed, n_customers=nc, max_order=max_order, max_stock=max_stock,
                                      num_steps=num_steps, env_offset=off)
        tag = (c, nc, max_order, max_stock, num_steps, E, masked)
        try:
            obs0, _ = env.reset_batch()
            kernels[env.exec_name] = kernels.get(env.exec_name, 0) + 1
            assert np.array_equal(obs0.cpu().numpy()[:, 0], v.reset()), "reset obs"
            for launch in range(int(r.randint(2, 5))):
                T = int(r.choice([1, 2, 3, 4, 5, 7, 8, 16, 33, 64, 100, 150]))
                A = r.uniform(-20, 1.3 * max_stock, size=(T, E, 1, 1)).astype(np.float32)
                M = (r.uniform(size=(T, E, 1)) > 0.15).astype(np.uint8) if masked else None
                host = launch == 1 and not masked
                if host:
                    ro = env.rollout_host(A)
                    obs, rew, ad = ro["observations"][:, :, 0], ro["rewards"][:, :, 0], ro["all_done"]
                else:
                    ro = env.rollout_batch(A, M)
                    obs = ro.observations.cpu().numpy()[:, :, 0]
                    rew = ro.rewards.cpu().numpy()[:, :, 0]
                    ad = ro.all_done.cpu().numpy()
                for t in range(T):
                    ref = v.step(A[t, :, 0, 0], None if M is None else M[t, :, 0])
                    want_obs = ref["obs"]
                    if ref["all_trunc"].all():
                        want_obs = v.reset()
                    assert np.array_equal(obs[t], want_obs), f"obs launch {launch} t {t}"
                    assert np.array_equal(rew[t], ref["reward"].astype(np.float32)), f"reward launch {launch} t {t}"
                    assert np.array_equal(ad[t, :, 1], ref["all_trunc"]), f"trunc launch {launch} t {t}"
                st = np.stack([np.atleast_1d(np.asarray(getattr(env.agents["SHOP"], k)))
                               for k in ("stock", "sales", "missed_sales", "delivered_stock")], axis=1)
                assert np.array_equal(st[:, 0], v.stock), f"stock after launch {launch}"
            env.check_errors()
        except Exception as exc:
            bad.append(tag)
            print("MISMATCH", tag, type(exc).__name__, str(exc)[:300], flush=True)
        env.close()
    print(json.dumps({"cases": a.count, "kernels": kernels, "mismatches": bad}))


if __name__ == "__main__":
    main()
