import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def port():
    """oracle/liboracle.so -- the CPU restatement (test infrastructure)."""
    from oracle import oracle as O
    O.build(ref=os.path.isdir("/root/reference"))
    return O.Port()


@pytest.fixture(scope="session")
def ref():
    """oracle/_ref/libmv_l2.so -- the unmodified reference, when it was built (it travels to the GPU box)."""
    from oracle import oracle as O
    if os.path.isdir("/root/reference"):
        O.build(ref=True)
    if not O.have_reference():
        pytest.skip("oracle/_ref/libmv_l2.so not built")
    return O.Reference()


@pytest.fixture(scope="session")
def libpath():
    from spmv_b200 import build
    return build.build()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
    return np.load(path)


def bits_equal(a, b):
    a = np.ascontiguousarray(a)
    b = np.ascontiguousarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a.view(np.uint8), b.view(np.uint8))
