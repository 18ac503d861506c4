#!/usr/bin/env python
"""How fast is the oracle port relative to the UNMODIFIED reference?  (build container only)

bench.py's CPU arm has to run on the GPU box, where /root/reference does not exist, so it
times `oracle.phantom_oracle` (kind "port").  This script times the same supply-chain episode
loop on the unmodified reference (through oracle/ref_shim.py, native np.random, telemetry at
its defaults, tracking off -- BASELINE.md section 3) and on the port, one core each, in the
same process on the same machine, and writes the ratio to tests/golden/port_calibration.json.
bench.py quotes that ratio next to its cpu_baseline so the reader can translate "port" speed
into reference speed.

    python tools/calibrate_port.py
"""
from __future__ import annotations

import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
T = 100


def rate(make_env, seconds=6.0):
    env = make_env()
    acts = np.random.RandomState(0).uniform(0, 100, size=(T, 1)).astype(np.float32)
    env.reset()
    for t in range(T):
        env.step({"SHOP": acts[t]})
    best = 0.0
    for _ in range(5):
        env.reset()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds / 5:
            for t in range(T):
                env.step({"SHOP": acts[t]})
            env.reset()
            n += T
        best = max(best, n / (time.perf_counter() - t0))
    return best


def main():
    from oracle import ref_shim, rng
    import oracle.phantom_oracle as po
    from oracle.workloads import supply_chain as wl

    sc = ref_shim.import_reference_supply_chain()
    ref = rate(lambda: sc.SupplyChainEnv())

    class NativeStream:  # the port's customers draw from np.random like the reference's
        def randint(self, n):
            return int(np.random.randint(n))

    port = rate(lambda: wl.build(po, NativeStream()))
    # and the port exactly as bench.py drives it (contract RNG stream + episode clock)
    out = {
        "workload": "supply-chain example, 1 env, 1 core, 100-step episodes incl. reset",
        "reference_env_steps_per_s": ref,
        "port_env_steps_per_s": port,
        "port_over_reference": port / ref,
        "host": os.popen("lscpu | grep 'Model name' | cut -d: -f2").read().strip(),
        "python": sys.version.split()[0],
        "numpy": np.__version__,
        "note": "best of 5 windows each; measured in the build container, same process",
    }
    path = os.path.join(REPO, "tests", "golden", "port_calibration.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
