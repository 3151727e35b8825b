import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture
def rng():
    # same seed as the reference's root conftest.py:28
    return np.random.default_rng(12345)


def pixie_like(n, C, seed=12345, nproto=30):
    """'P' distribution of SURVEY.md section 8d: Dirichlet prototypes + noise, row-normalised,
    then divided by the per-channel 99.9th percentile (what Pixie preprocessing produces)."""
    r = np.random.default_rng(seed)
    protos = r.dirichlet(np.full(C, 0.3), size=nproto)
    which = r.integers(0, nproto, n)
    X = protos[which] + np.abs(r.normal(0, 0.05, (n, C)))
    X = np.maximum(X, 0)
    X /= X.sum(1, keepdims=True)
    X /= np.quantile(X, 0.999, axis=0)
    return X.astype(np.float32)
