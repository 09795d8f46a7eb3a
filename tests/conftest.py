import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    """CPU oracles (test infrastructure): the C restatement is built on demand, the reference build is used if present."""
    import oracle

    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def shf():
    import superterrainplus_b200 as pkg

    return pkg
