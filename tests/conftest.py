import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def native():
    from zksnake_b200 import _native
    return _native


@pytest.fixture(scope="session")
def gpu(native):
    if not native.gpu_available():
        pytest.fail("no CUDA device visible: -m gpu tests need a B200 (there is no CPU fallback)")
    native.ensure_init()
    return native
