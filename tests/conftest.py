import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def lib():
    from miniraytracer_b200 import api
    return api.load()


@pytest.fixture(scope="session")
def emul_bin():
    """TEST-ONLY g++ build of the device tracer core (tests/host_emul)."""
    import oracle_util
    return oracle_util.build_emul()


@pytest.fixture(scope="session")
def emul_kernel_bin():
    """TEST-ONLY g++ build of the render kernels themselves behind a SIMT shim (tests/host_emul/emul_warp.h)."""
    import oracle_util
    return oracle_util.build_emul_binned()
