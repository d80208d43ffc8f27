import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def lj_sample():
    return dict(np.load(os.path.join(GOLDEN, "lj_sample.npz")))


@pytest.fixture(scope="session")
def dioxin_water():
    return dict(np.load(os.path.join(GOLDEN, "dioxin_water.npz")))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_c

    oracle_c.build()
    return oracle_c


@pytest.fixture(scope="session")
def em():
    import emdee_jl_b200

    return emdee_jl_b200
