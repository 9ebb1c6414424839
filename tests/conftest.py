import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)


def rel_l2_mag(a, b):
    """|| |a| - |b| ||_2 / || |b| ||_2  (the north-star parity metric)."""
    a, b = np.abs(np.asarray(a)), np.abs(np.asarray(b))
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def rel_l2(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
