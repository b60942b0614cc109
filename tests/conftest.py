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


@pytest.fixture(scope="session")
def ops_golden():
    return dict(np.load(os.path.join(GOLDEN, "ops_golden.npz")))


@pytest.fixture(scope="session")
def train_golden():
    return dict(np.load(os.path.join(GOLDEN, "train_golden.npz")))


def rel_err(a, ref):
    """Normwise error of SURVEY §8d: max|a-ref| / max|ref| (NaN positions must coincide)."""
    a = np.asarray(a, np.float64)
    ref = np.asarray(ref, np.float64)
    assert a.shape == ref.shape, (a.shape, ref.shape)
    nan_a, nan_r = np.isnan(a), np.isnan(ref)
    assert np.array_equal(nan_a, nan_r), "NaN pattern differs"
    fin = ~nan_r & np.isfinite(ref)
    if not fin.any():
        return 0.0
    denom = max(np.abs(ref[fin]).max(), 1e-30)
    return float(np.abs(a[fin] - ref[fin]).max() / denom)
