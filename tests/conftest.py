import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", category=DeprecationWarning)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def rel_err(a, b, floor=1e-6):
    """Tensor-level relative error max|a-b| / max(floor, max|b|) — the tolerance metric used by every
    parity test (rel 1e-5 for the FP32 path, 2e-2 for the TF32 tensor-core path).  ``floor`` keeps
    tensors that are mathematically zero (e.g. the bias gradient in front of a BatchNorm, ~1e-18 of
    round-off in the float64 reference) from dividing noise by noise."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    if np.isnan(a).any() or np.isnan(b).any():
        assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
        a, b = np.nan_to_num(a), np.nan_to_num(b)
    denom = max(float(np.max(np.abs(b))), floor)
    return float(np.max(np.abs(a - b))) / denom


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


@pytest.fixture(scope="session")
def ns():
    from tests import topologies
    return topologies.product_ns()


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
