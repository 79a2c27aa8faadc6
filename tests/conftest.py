import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """the CPU checkers (oracle/): plain-C restatement + the reference's own object code"""
    from oracle import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def dfb():
    """the product: ctypes mirror over libdfb200.so (hard error if the library is not built)"""
    import _dfb_import  # noqa: F401
    import digital_filtering_b200
    digital_filtering_b200.lib()
    return digital_filtering_b200


@pytest.fixture(scope="session")
def W():
    import _dfb_import  # noqa: F401
    from digital_filtering_b200 import workloads
    return workloads


def normwise_close(got, ref, tol=1e-12):
    """north_star's fp64 gate, in the normwise form SURVEY section 7 shows is the attainable one:
    |got - ref| <= tol * max(|ref|, rms(ref)) elementwise.  Returns (ok, worst ratio)."""
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    rms = float(np.sqrt(np.mean(ref * ref)))
    scale = np.maximum(np.abs(ref), rms if rms > 0 else 1.0)
    ratio = float(np.max(np.abs(got - ref) / scale)) if ref.size else 0.0
    return ratio <= tol, ratio
