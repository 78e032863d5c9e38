import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_random():
    with open(os.path.join(GOLDEN_DIR, "random_circuits.json")) as fh:
        return json.load(fh)["cases"]


@pytest.fixture(scope="session")
def golden_shipped():
    with open(os.path.join(GOLDEN_DIR, "shipped_circuits.json")) as fh:
        return json.load(fh)["circuits"]


@pytest.fixture(scope="session")
def golden_large_primes():
    with open(os.path.join(GOLDEN_DIR, "large_primes.json")) as fh:
        return json.load(fh)["cases"]
