import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _have_gpu():
    """True if the C ABI finds an sm_100 device (mrt_gpu_init), without importing torch."""
    try:
        from miniraytracer_b200 import api
        return api.load(build_if_missing=False).mrt_gpu_init(0, None) == 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `gpu` tests are the parity tests proper and need a B200: skip (not error) on a box without one, unless the
    # run asked for them explicitly with -m gpu -- there a missing device / library must fail loudly
    if "gpu" in (config.getoption("-m") or "") and "not gpu" not in (config.getoption("-m") or ""):
        return
    if _have_gpu():
        return
    skip = pytest.mark.skip(reason="no sm_100 CUDA device on this box")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


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
