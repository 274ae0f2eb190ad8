import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def golden_inner():
    """basisFuncsInner with caller-chosen indices, from the reference's own compiled C++
    (tests/golden/gen_basisfuncs_inner_golden.py)."""
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "basisfuncs_inner_reference.npz"))


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bspline_reference.npz"))
