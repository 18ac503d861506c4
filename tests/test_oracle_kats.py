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
