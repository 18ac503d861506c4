"""The reference's KATs (restated in tests/kat_scenarios.py) against the CPU oracle."""
import pytest

import oracle.phantom_oracle as po
from oracle.workloads import mock

from . import kat_scenarios as kats


@pytest.fixture(scope="module")
def K():
    return mock.build_classes(po)


@pytest.mark.parametrize("scenario", kats.ALL, ids=lambda f: f.__name__)
def test_kat_on_oracle(K, scenario):
    scenario(K)


def test_handler_kats_on_the_unmodified_reference():
    """The three handler-driven FSM scenarios (fsm.py:294-307; two restate the reference's own
    tests, one is authored) executed by the UNMODIFIED reference, imported through
    oracle/ref_shim.py in a fresh interpreter: the expectations hard-coded in
    tests/kat_scenarios.py (FSM_STATE_DRIVEN_TRACE ...) are the reference's, not the port's.
    Only possible where /root/reference exists (the build container)."""
    import os
    import subprocess
    import sys

    from oracle import ref_shim

    if not os.path.isdir(os.path.join(ref_shim.REFERENCE_ROOT, "phantom")):
        pytest.skip("the reference is only available in the build container")
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "from oracle import ref_shim\n"
        "from oracle.workloads import mock\n"
        "from tests import kat_scenarios as k\n"
        "K = mock.build_classes(ref_shim.import_reference())\n"
        "for f in (k.scenario_fsm_one_state_with_handler, k.scenario_fsm_invalid_transition_runtime,\n"
        "          k.scenario_fsm_handler_state_driven, k.scenario_fsm_one_state,\n"
        "          k.scenario_fsm_odd_even_two_agents, k.scenario_stackelberg_acting_order):\n"
        "    f(K)\n"
        "print('reference ok')\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=repo, capture_output=True, text=True,
                         timeout=300)
    assert out.returncode == 0 and "reference ok" in out.stdout, out.stderr[-2000:]


def test_random_handler_fsms_match_the_reference(K):
    """Differential fuzz of the FSM env-handler path (fsm.py:294-307): 40 random
    FiniteStateMachineEnvs (tests/kat_scenarios.py:random_handler_fsm) on the oracle port ==
    the traces the UNMODIFIED reference produced for the same case seeds
    (tests/golden/fsm_handler_fuzz_reference.json, oracle/make_golden.py)."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "fsm_handler_fuzz_reference.json")
    want = json.load(open(path))
    assert len(want) == 40 and sum(t[-1][0] == "raise" for t in want.values()) >= 2
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s)))
        assert got == want[str(s)], f"case seed {s}"
    # compound handlers: if / elif / else chains over the clock, agent counters and constants
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_compound_fuzz")))
    assert len(want) == 40 and sum(t[-1][0] == "raise" for t in want.values()) >= 2
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, compound=True)))
        assert got == want[str(s)], f"compound case seed {s}"
    # float32 comparisons (an echo agent's float32 `level`) next to the integer ones
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_float_fuzz")))
    assert len(want) == 32
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, floats=True)))
        assert got == want[str(s)], f"float case seed {s}"
    # cases whose outcome depends on the order of FSMStage.acting_agents (fsm.py:276-277)
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_order_fuzz")))
    assert len(want) == 13
    for s in want:
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, int(s), floats=True)))
        assert got == want[s], f"order case seed {s}"
    # mail that waits across steps (handlers that do not resolve)
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_waiting_fuzz")))
    assert len(want) == 40
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, waiting=True)))
        assert got == want[str(s)], f"waiting case seed {s}"
    # env classes wider than a warp: 33..120 agents
    want = json.load(open(path.replace("fsm_handler_fuzz", "fsm_wide_fuzz")))
    assert len(want) == 16
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_random_handler_fsm(K, s, wide=True)))
        assert got == want[str(s)], f"wide case seed {s}"


def test_random_base_and_stackelberg_envs_match_the_reference(K):
    """60 random PhantomEnv / StackelbergEnv env classes over the mock agents (halving and
    request / response echoes, random graphs, ignore_connection_errors, round limits, receivers
    without a handler, agents terminating mid-episode, leader / follower lists in random order,
    shuffled batches) on the oracle port == the traces of the UNMODIFIED reference, message lists
    and exception types included (tests/golden/mock_env_fuzz_reference.json)."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden",
                        "mock_env_fuzz_reference.json")
    want = json.load(open(path))
    assert len(want) == 60 and sum(t[-1][0] == "raise" for t in want.values()) >= 3
    assert sum(bool(t[0][1]) for t in want.values()) >= 10  # shuffled cases
    for s in range(len(want)):
        got = json.loads(json.dumps(kats.run_mock_env(K, s)))
        assert got == want[str(s)], f"case seed {s}"
