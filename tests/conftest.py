import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def has_gpu():
    try:
        from orphics_b200 import _capi
        _capi.require_device()
        return True
    except Exception:
        return False


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def theory():
    from oracle import theory as oth
    return oth.load_theory()


def relerr(a, b):
    """max |a-b| / max |b| (the SURVEY 8d convention for maps)."""
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))
