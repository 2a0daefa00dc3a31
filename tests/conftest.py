import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_golden.npz"), allow_pickle=False)


@pytest.fixture(scope="session")
def contracts():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "state_dict_contracts.json")) as f:
        return json.load(f)
