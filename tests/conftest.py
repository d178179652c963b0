import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def gpu_api_cls():
    """The product backend.  Fails loudly (no CPU fallback) when the CUDA library or the device is missing."""
    from horses3d_b200.capi import GpuApi
    return GpuApi
