import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA (B200) device; run with -m gpu")
    config.addinivalue_line("markers", "slow: longer CPU test (still part of the default CPU suite)")


def has_gpu():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a B200 skips the GPU tests instead of failing them."""
    if has_gpu():
        return
    skip = pytest.mark.skip(reason="needs a B200 (no CUDA device visible)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def taipei():
    from dsurftomo_b200 import inputs

    return inputs.config(1)


@pytest.fixture(scope="session")
def small_problem():
    """A small heterogeneous synthetic problem with all four data types (fast on the oracle)."""
    from dsurftomo_b200 import inputs

    return inputs.synthetic_problem(12, 3, 6, ("Rc", "Rg", "Lc", "Lg"), nrecv=5, name="small_4types")
