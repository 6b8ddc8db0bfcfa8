import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    # a fresh checkout has no built library (it is git-ignored) and the package refuses to import without it
    if not os.path.exists(os.path.join(ROOT, "synthpy_b200", "csrc", "libsynthpy_b200.so")):
        import __graft_entry__
        __graft_entry__.build()


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]
    return load


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|, floor) with NaN positions required to coincide."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs"
    m = ~na
    if not m.any():
        return 0.0
    scale = np.maximum(np.abs(b[m]), floor) if floor else np.abs(b[m])
    scale = np.where(scale == 0, 1.0, scale)
    return float(np.max(np.abs(a[m] - b[m]) / scale))
