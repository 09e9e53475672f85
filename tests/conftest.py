import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as fh:
        return json.load(fh)


@pytest.fixture(scope="session")
def golden_arrays():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    return load


def sample_idx(numel, k=256):
    g = np.random.RandomState(12345)
    return np.sort(g.choice(numel, size=min(k, numel), replace=False))
