#!/usr/bin/env python
"""One-off campaign 3 (GPU box): (a) tests/test_gpu_fuzz.py's random supply-chain topologies for
many more seeds; (b) random PhantomEnv / StackelbergEnv env classes over the mock agents --
halving and request/response echo agents, random graphs, ignore_connection_errors, round limits
(RuntimeError), sends without an edge (NetworkError), agents that terminate mid-episode, leader /
follower lists in random order, message tracking -- on every tiling against the CPU oracle port:
observations, rewards, done flags, call counters, float32 levels, exception types AND the
tracked message list of every step.

    python tools/fuzz_campaign3.py [--first 0] [--count 1000] [--chains 300]
"""
import argparse
import json
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

from tests.kat_scenarios import counts  # noqa: E402  (before anything can shadow `tests`)


from tests.kat_scenarios import random_mock_env, run_mock_env  # noqa: E402,F401


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=1000)
    ap.add_argument("--chains", type=int, default=300)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle.workloads import mock as omock
    from tools.fuzz_campaign import device_ns

    KO, KD = omock.build_classes(po), device_ns()
    bad = []
    for s in range(a.first, a.first + a.count):
        want = json.loads(json.dumps(run_mock_env(KO, s)))
        # (shuffled batches need the per-receiver lists of the tile / block engines)
        for mode in (("queue", "wide") if want[0][1] else ("thread", "queue", "wide")):
            KD.ph.PhantomEnv.default_exec_mode = mode
            try:
                got = json.loads(json.dumps(run_mock_env(KD, s)))
            except Exception as exc:
                got = ["exception", type(exc).__name__, str(exc)[:200]]
            if got != want:
                bad.append((s, mode))
                k = next((i for i, (x, y) in enumerate(zip(got, want)) if x != y), -1)
                print("MISMATCH mock", s, mode, "first differing record", k, str(got[k] if 0 <= k < len(got) else got)[:400],
                      "WANT", str(want[k] if 0 <= k < len(want) else want)[:400], flush=True)
    KD.ph.PhantomEnv.default_exec_mode = "auto"
    print("mock", {"cases": a.count, "mismatches": bad}, flush=True)
    chain_bad = []
    import tests.test_gpu_fuzz as tf

    for s in range(100, 100 + a.chains):
        for mode in ("thread", "queue"):
            try:
                tf.test_random_supply_chain_topology(s, mode)
            except Exception as exc:
                chain_bad.append((s, mode))
                print("MISMATCH chain", s, mode, type(exc).__name__, str(exc)[:300], flush=True)
    print(json.dumps({"mock_cases": a.count, "mock_mismatches": bad, "chain_cases": a.chains,
                      "chain_mismatches": chain_bad}))


if __name__ == "__main__":
    main()
