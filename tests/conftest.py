import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "slow: larger CPU cases")


@pytest.fixture(scope="session")
def tiny_assets():
    from supertonic_b200 import surrogate
    return surrogate.ensure_assets("tiny")


@pytest.fixture(scope="session")
def full_assets():
    from supertonic_b200 import surrogate
    return surrogate.ensure_assets("full")


@pytest.fixture(scope="session")
def host_golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "host_golden.json"), encoding="utf-8") as f:
        return json.load(f)["results"]
