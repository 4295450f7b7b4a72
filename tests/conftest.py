import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "reference_outputs.npz"))


@pytest.fixture(scope="session")
def clouds():
    return np.load(os.path.join(GOLDEN_DIR, "tutorial_clouds.npz"))


CLOUD_NAMES = ["Boxy_smooth100k", "0000000000"]
