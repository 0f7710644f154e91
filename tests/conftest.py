import json
import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
os.environ.setdefault("TZ", "UTC")
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def chain_golden():
    return load_golden("chain_kats.json")


@pytest.fixture(scope="session")
def gpu():
    """Bind the CUDA device through the C ABI; fails (does not skip) when there is none."""
    from fei_b200 import _abi
    _abi.init()
    return _abi
