#!/usr/bin/env python
"""One-off campaign 6 (GPU box): RUN-TIME SPECIALISED units (env.specialise(): the step kernel
rebuilt by nvcc with the env class as a compile-time constant) of random handler-driven FSM env
classes -- compound / float32 rules, custom acting orders, waiting mail -- on the thread-per-env
and the tile engine, against the CPU oracle port.  One nvcc run per case and engine.

    python tools/fuzz_campaign6.py [--first 0] [--count 40]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tools"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=0)
    ap.add_argument("--count", type=int, default=40)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle.workloads import mock as omock
    from fuzz_campaign import device_ns
    from tests import kat_scenarios as kats

    KO, KD = omock.build_classes(po), device_ns()
    names, bad, runs = {}, [], 0

    def prepare(env):
        env.specialise()
        names[env.exec_name] = names.get(env.exec_name, 0) + 1

    for variant, kw, modes in (("compound", {"compound": True}, ("thread", "queue")),
                               ("float32", {"floats": True}, ("thread", "queue")),
                               ("waiting-mail", {"waiting": True}, ("thread",))):
        for s in range(a.first, a.first + a.count):
            want = json.loads(json.dumps(kats.run_random_handler_fsm(KO, s, **kw)))
            for mode in modes:
                KD.ph.PhantomEnv.default_exec_mode = mode
                runs += 1
                try:
                    got = json.loads(json.dumps(kats.run_random_handler_fsm(KD, s, prepare=prepare, **kw)))
                except Exception as exc:
                    got = ["exception", type(exc).__name__, str(exc)[:200]]
                if got != want:
                    bad.append((variant, s, mode))
                    print("MISMATCH", variant, s, mode, str(got)[:300], flush=True)
    KD.ph.PhantomEnv.default_exec_mode = "auto"
    print(json.dumps({"runs": runs, "kernels": names, "mismatches": bad}))


if __name__ == "__main__":
    main()
