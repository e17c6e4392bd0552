import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def xt():
    """The expression mirror with the CPU oracle registered as evaluator of HostArray trees."""
    from oracle import oracle
    return oracle.install()


@pytest.fixture(scope="session")
def gpu(xt):
    """Initialise device 0; GPU tests fail loudly (not skip) if the extension or device is missing."""
    from xtensor_b200 import capi
    capi.check(capi.lib().xtb_init(0))
    return capi.lib()
