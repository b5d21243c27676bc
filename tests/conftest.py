"""pytest configuration: path setup, the `gpu` marker, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "pytorch-deepfepe_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden():
    """Outputs of the unmodified reference on seeded inputs (tests/golden/make_golden.py)."""
    path = os.path.join(ROOT, "tests", "golden", "reference_outputs.npz")
    return dict(np.load(path, allow_pickle=False))
