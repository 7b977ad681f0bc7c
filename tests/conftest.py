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
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout (build container only)")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


def assert_close_t1(a, b, rtol=1e-5, atol=1e-6, what=""):
    """T1 tier of SURVEY.md section 8: 1e-5 relative in FP32 (atol covers cancellation near 0)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), "%s: max violation %g at %s" % (what, err.max(), np.unravel_index(err.argmax(), err.shape))
