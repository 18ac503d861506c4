#!/usr/bin/env python
"""One-off differential campaign (not part of the test suite): random handler-driven FSM env
classes (tests/kat_scenarios.py:random_handler_fsm -- plain / compound / float32 / wide) on the
device, every engine tiling, against the CPU oracle port (which the reference's own tests and the
committed fuzz fixtures pin).  Prints one line per variant and every mismatching case seed.

    python tools/fuzz_campaign.py [--first 100] [--count 150]
"""
import argparse
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def device_ns():
    import phantom_b200 as ph
    from phantom_b200.envs import mock

    class NS:
        pass

    ns = NS()
    ns.ph = ph
    ns.MockAgent, ns.MockStrategicAgent, ns.EchoAgent = mock.MockAgent, mock.MockStrategicAgent, mock.EchoAgent
    ns.finish_network = lambda network: network
    ns.CodecAgent = mock.CodecAgent
    ns.stage_handler = ph.StageRule
    return ns


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--first", type=int, default=100)
    ap.add_argument("--count", type=int, default=150)
    a = ap.parse_args()
    import oracle.phantom_oracle as po
    from oracle.workloads import mock as omock
    from tests import kat_scenarios as kats

    KO, KD = omock.build_classes(po), device_ns()
    report = {}
    for variant, kw, modes, count in (("plain", {}, ("thread", "queue", "wide"), a.count),
                                      ("compound", {"compound": True}, ("thread", "queue", "wide"), a.count),
                                      ("float32", {"floats": True}, ("thread", "queue", "wide"), a.count),
                                      ("waiting-mail", {"waiting": True}, ("thread",), a.count),
                                      ("wide", {"wide": True}, ("auto",), max(a.count // 5, 1))):
        bad = []
        for s in range(a.first, a.first + count):
            want = json.loads(json.dumps(kats.run_random_handler_fsm(KO, s, **kw)))
            for mode in modes:
                KD.ph.PhantomEnv.default_exec_mode = mode
                try:
                    got = json.loads(json.dumps(kats.run_random_handler_fsm(KD, s, **kw)))
                except Exception as exc:  # a create-time refusal is a finding too
                    got = ["exception", type(exc).__name__, str(exc)[:200]]
                if got != want:
                    bad.append((s, mode))
                    print("MISMATCH", variant, s, mode, str(got)[:300], flush=True)
        report[variant] = {"cases": count, "modes": list(modes), "mismatches": bad}
        print(variant, report[variant], flush=True)
    KD.ph.PhantomEnv.default_exec_mode = "auto"
    print(json.dumps(report))


if __name__ == "__main__":
    main()
