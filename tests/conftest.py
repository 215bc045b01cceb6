import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CPU oracle is test infrastructure: make sure it is compiled. The CUDA library is built by
    __graft_entry__.build(); GPU tests fail loudly if it is missing (no fallback)."""
    from oracle import lporacle
    lporacle.build()
    yield
