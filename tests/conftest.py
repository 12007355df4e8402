import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(autouse=True, scope="session")
def _single_threaded_oracle():
    """The CPU oracle reproduces the real reference bit for bit at ONE thread (tests/test_oracle_golden.py); multi-threaded
    mkldnn convolutions alone move disparities by ~1e-4 px and flip rounding-level top-2 ties, which is the
    size of what the GPU parity tests measure.  Every test therefore runs the oracle single-threaded."""
    import torch
    torch.set_num_threads(1)
    yield
