"""TEST INFRASTRUCTURE.  Pins the oracle against the reference's own tests.

Runs the reference's test files for the env-step hot path twice:
  1. against the unmodified reference imported through oracle/ref_shim.py
     (proves the third-party stubs are faithful), and
  2. against oracle.phantom_oracle registered under the name `phantom`
     (proves the restatement reproduces every golden vector / KAT the reference holds).

Only usable in the build container (needs /root/reference).  The result is recorded in
tests/golden/PINNED.json, which tests/test_oracle_pinned.py checks on every run; the
KAT vectors themselves are also restated in tests/test_oracle_kats.py so that they run
where /root/reference does not exist.

    python -m oracle.pin_against_reference            # run + write PINNED.json
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import subprocess
import sys

from . import ref_shim

HOT_PATH_TESTS = [
    "tests/test_env.py",
    "tests/test_agent.py",
    "tests/test_message.py",
    "tests/test_stackelberg.py",
    "tests/network/test_network.py",
    "tests/network/test_resolver.py",
    "tests/network/test_tracking.py",
    "tests/network/test_payload_checks.py",
    "tests/network/test_stochastic_network.py",
    "tests/fsm/test_fsm_init.py",
    "tests/fsm/test_fsm_validation.py",
    "tests/fsm/test_is_fsm_deterministic.py",
    "tests/fsm/test_odd_even_one_agent.py",
    "tests/fsm/test_odd_even_two_agents.py",
    "tests/fsm/test_one_state.py",
    "tests/encoders/test_chained.py",
    "tests/encoders/test_dict.py",
    "tests/decoders/test_chained.py",
    "tests/decoders/test_dict.py",
    "tests/test_supertypes_env.py",
]

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(target: str) -> dict:
    env = dict(os.environ, PHX_PIN_TARGET=target, PYTHONHASHSEED="1",
               PYTHONDONTWRITEBYTECODE="1",
               PYTHONPATH=REPO + os.pathsep + os.environ.get("PYTHONPATH", ""))
    cmd = [sys.executable, "-m", "pytest", "-q", "-p", "no:cacheprovider",
           "-p", "oracle._pin_plugin", "--rootdir", ref_shim.REFERENCE_ROOT,
           "-c", "/dev/null"]
    cmd += [os.path.join(ref_shim.REFERENCE_ROOT, t) for t in HOT_PATH_TESTS]
    p = subprocess.run(cmd, cwd="/tmp", env=env, capture_output=True, text=True)
    tail = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else p.stderr[-400:]
    counts = {k: int(v) for v, k in re.findall(r"(\d+) (passed|failed|error|errors|skipped)", tail)}
    return {"target": target, "returncode": p.returncode, "summary": tail,
            "counts": counts, "stdout": p.stdout, "stderr": p.stderr}


def reference_fingerprint() -> str:
    h = hashlib.sha256()
    for t in HOT_PATH_TESTS:
        with open(os.path.join(ref_shim.REFERENCE_ROOT, t), "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def oracle_fingerprint() -> str:
    h = hashlib.sha256()
    d = os.path.join(REPO, "oracle", "phantom_oracle")
    for name in sorted(os.listdir(d)):
        if name.endswith(".py"):
            with open(os.path.join(d, name), "rb") as f:
                h.update(name.encode() + f.read())
    return h.hexdigest()


def main() -> int:
    if not ref_shim.reference_available():
        print("reference tree not available; nothing to pin against")
        return 2
    out = {"reference_commit": "9ce42f2 (v2.2.0)", "tests": HOT_PATH_TESTS,
           "reference_tests_sha256": reference_fingerprint(),
           "oracle_sha256": oracle_fingerprint()}
    rc = 0
    for target in ("reference", "oracle"):
        r = run(target)
        print(f"[{target}] {r['summary']}")
        if r["returncode"] != 0:
            rc = 1
            print(r["stdout"][-6000:])
            print(r["stderr"][-2000:])
        out[target] = {"returncode": r["returncode"], "summary": r["summary"],
                       "counts": r["counts"]}
    dst = os.path.join(REPO, "tests", "golden", "PINNED.json")
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    with open(dst, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", dst)
    return rc


if __name__ == "__main__":
    sys.exit(main())
