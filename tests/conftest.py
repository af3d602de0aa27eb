import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "gpu_unverified: needs a B200 AND exercises code that has not run on one yet "
                                       "(opt in with COZO_RUN_UNVERIFIED=1; never selected by -m gpu)")


@pytest.fixture(scope="session")
def gpu():
    """Initialise the CUDA library; fails loudly (no CPU fallback) when unusable."""
    from cozo_b200 import capi
    capi.init(0)
    return capi
