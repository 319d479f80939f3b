import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_devices() -> int:
    try:
        from smm_jl_b200 import _lib
        return int(_lib.device_count())
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: without one they are skipped (the product has no CPU fallback to test)."""
    if not any("gpu" in item.keywords for item in items):
        return
    if _cuda_devices() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the CUDA path cannot run here (and there is no CPU fallback)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure); built on demand with g++."""
    from oracle import oracle_lib
    oracle_lib.lib()
    return oracle_lib


@pytest.fixture(scope="session")
def smm():
    """The product's C-ABI binding; requires a CUDA device for compute calls."""
    from smm_jl_b200 import _lib
    _lib.lib()
    return _lib
