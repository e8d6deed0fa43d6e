import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def env_golden():
    return np.load(os.path.join(GOLDEN, "env_golden.npz"))


@pytest.fixture(scope="session")
def env27_golden():
    return np.load(os.path.join(GOLDEN, "env27_golden.npz"))


@pytest.fixture(scope="session")
def dqn_golden():
    return np.load(os.path.join(GOLDEN, "dqn_golden.npz"))


def episode(g, i):
    """Unpack episode i of a golden file into a dict."""
    pre = "ep%d_" % i
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
