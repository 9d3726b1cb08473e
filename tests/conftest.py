import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def example_golden():
    return dict(np.load(os.path.join(GOLDEN, "example_pair.npz")))


@pytest.fixture(scope="session")
def small_golden():
    return dict(np.load(os.path.join(GOLDEN, "small_pair.npz")))


@pytest.fixture(scope="session")
def nsga_golden():
    return dict(np.load(os.path.join(GOLDEN, "nsga2.npz")))


@pytest.fixture(scope="session")
def edits_golden():
    return dict(np.load(os.path.join(GOLDEN, "edits.npz")))


@pytest.fixture(scope="session")
def mt_golden():
    return dict(np.load(os.path.join(GOLDEN, "mt_stream.npz")))


def dense_from_coo(n1, n2, r, c, v):
    S = np.zeros((n1, n2), np.float64)
    S[r, c] = v
    return S
