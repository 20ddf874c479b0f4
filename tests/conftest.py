import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: longer CPU-side check")


@pytest.fixture(scope="session")
def goldens():
    import numpy as np
    d = os.path.join(ROOT, "tests", "golden")
    return dict(np.load(os.path.join(d, "gmg_goldens.npz"))), dict(np.load(os.path.join(d, "gmg_ref_ops.npz")))


@pytest.fixture(scope="session")
def gmg_oracle():
    import oracle
    return oracle.gmg()


@pytest.fixture(scope="session")
def gmg_ref():
    import oracle
    r = oracle.ref_gmg()
    if r is None:
        pytest.skip("oracle/_ref not built (no /root/reference here); goldens still pin the oracle")
    r.set_threads(1)
    return r
