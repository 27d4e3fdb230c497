import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def O():
    """The parity oracle (test infrastructure; see oracle/aukit_oracle.h)."""
    from oracle import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def ak():
    """The product package, with its CUDA library built."""
    from aukit_b200 import build
    build.build()
    import aukit_b200
    return aukit_b200
