#!/usr/bin/env python
"""One-off campaign (build container, CPU): the ORACLE PORT against the UNMODIFIED reference on the
random env-class generators of tests/kat_scenarios.py -- the same generators the device is
compared with the oracle on (tools/fuzz_campaign*.py).  Closes the chain device == oracle ==
reference for those campaigns.  Needs /root/reference.

    python tools/oracle_vs_reference_campaign.py
"""
import sys, json, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import kat_scenarios as kats
import oracle.phantom_oracle as po
from oracle import ref_shim
from oracle.workloads import mock as omock

if __name__ == "__main__":
    KO = omock.build_classes(po)
    KR = omock.build_classes(ref_shim.import_reference())
    rep = {}
    t0 = time.time()
    for name, fn, n in (("plain", lambda K, s: kats.run_random_handler_fsm(K, s), 10000),
                        ("compound", lambda K, s: kats.run_random_handler_fsm(K, s, compound=True), 10000),
                        ("float32", lambda K, s: kats.run_random_handler_fsm(K, s, floats=True), 10000),
                        ("waiting-mail", lambda K, s: kats.run_random_handler_fsm(K, s, waiting=True), 10000),
                        ("wide", lambda K, s: kats.run_random_handler_fsm(K, s, wide=True), 500),
                        ("mock base/stackelberg", lambda K, s: kats.run_mock_env(K, s), 10000)):
        bad = []
        for s in range(10000, 10000 + n):
            a = json.loads(json.dumps(fn(KO, s)))
            b = json.loads(json.dumps(fn(KR, s)))
            if a != b:
                bad.append(s)
        rep[name] = {"cases": n, "mismatches": bad}
        print(name, rep[name], round(time.time() - t0), "s", flush=True)
    print(json.dumps(rep))
