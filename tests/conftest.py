import os
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def checker():
    import harness
    return harness.checker_lib()


@pytest.fixture(scope="session")
def ref():
    import harness
    lib = harness.ref_lib()
    if lib is None:
        pytest.skip("compiled reference (oracle/_ref) not available")
    return lib
