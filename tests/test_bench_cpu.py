"""bench.py's host-side helpers (no GPU): the committed evidence files it quotes exist and parse,
the algorithmic-byte table matches DESIGN.md's per-config figures, and the reference arm's CPU
leg steps the unmodified reference when baseline/_ref is present (else the oracle port)."""
import importlib
import json
import os

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    import sys

    sys.path.insert(0, REPO)
    return importlib.import_module("bench")


def test_traffic_and_ceiling_records(bench):
    total, src = bench.ncu_traffic()
    assert total and 5e7 < total < 2e8 and "profiles/" in src
    assert os.path.exists(os.path.join(REPO, src.split(" ")[0]))
    for cfg in ("C4", "C5"):
        t, s = bench.ncu_traffic(cfg)
        assert t and t > 1e8 and os.path.exists(os.path.join(REPO, s.split(" ")[0]))
    assert bench.ncu_traffic("C3") == (None, None)  # no capture committed for that kernel
    c1 = bench.pcie_ceiling(1, 2.7e9)
    assert 0.8 < c1["frac_of_ceiling"] < 1.0 and c1["ceiling"] > 2.5e9
    c8 = bench.pcie_ceiling(8, 5.0e9)
    assert c8["ceiling"] > c1["ceiling"] and 0.8 < c8["frac_of_ceiling"] <= 1.0
    assert bench.pcie_ceiling(3, 1.0) == {}


def test_algorithmic_bytes_table(bench):
    """bytes per launch = E * (T * b_io + b_state), DESIGN.md section 3 / SURVEY 8(d)."""
    c = bench.CONFIGS
    assert set(c) == {"C2", "C3", "C4", "C5"}
    assert c["C2"]["E"] * (c["C2"]["T"] * c["C2"]["b_io"] + c["C2"]["b_state"]) == 147324928
    assert (c["C3"]["E"], c["C3"]["S"], c["C4"]["E"], c["C5"]["E"], c["C5"]["S"]) == (32768, 31, 131072, 16384, 128)
    assert c["C4"]["scaling"] == c["C5"]["scaling"] == "strong" and c["C2"]["scaling"] == "weak"
    peaks = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(peaks):
        assert json.load(open(peaks))["hbm_gbs"] > 1000


def test_cpu_arm_kind(bench):
    kind = bench.cpu_kind()
    have_ref = os.path.isdir(os.path.join(REPO, "baseline", "_ref", "phantom"))
    assert kind == ("reference" if have_ref else "port")
