import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def bp():
    import bulletproofs_amcl_b200 as m
    return m


@pytest.fixture(scope="session")
def ctx_bls(bp):
    c = bp.Context(bp.BLS12_381, 0)
    yield c
    c.close()


@pytest.fixture(scope="session")
def ctx_bn(bp):
    c = bp.Context(bp.BN254, 0)
    yield c
    c.close()
