import os
import sys

import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


@pytest.fixture(scope="session")
def pkg():
    import xgtest
    return xgtest.package()


@pytest.fixture(scope="session")
def reflib():
    import xgtest
    L = xgtest.ref_lib()
    if L is None:
        pytest.skip("oracle/_ref/libfrenc_ref.so not built (no /root/reference here)")
    return L
