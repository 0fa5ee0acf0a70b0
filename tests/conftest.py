import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU must fail loudly, not skip silently: the product has no CPU path.
    pass


@pytest.fixture(scope="session")
def oracle():
    from oracle import usrt_oracle
    usrt_oracle.build()
    return usrt_oracle


@pytest.fixture(scope="session")
def usrt():
    """The product package, with the CUDA library loaded (fails loudly if it is not built)."""
    from unitysimpleraytracing_b200 import _lib, host
    _lib.load()
    return host
