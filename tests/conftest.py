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
    """The product library and the oracle port are built in-tree (no-ops when up to date)."""
    from smcpp_b200 import capi
    from oracle import obsport, port
    capi.build()
    port.build()
    obsport.build()
