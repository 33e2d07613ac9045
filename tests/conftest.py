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


@pytest.fixture
def rng():
    # same seed as the reference's tests/conftest.py:20-22
    return np.random.default_rng(50)


@pytest.fixture(scope="session")
def oracle():
    import oracle as O

    O.build()
    return O


@pytest.fixture(scope="session")
def hb():
    import heracles_b200

    return heracles_b200


@pytest.fixture(scope="session")
def ctx(hb):
    return hb.get_context(0)


def golden(name):
    return np.load(os.path.join(GOLDEN, name))
