import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.lib()
    return oracle


@pytest.fixture(scope="session")
def cuda_lib():
    """libscda_b200 through its C ABI; GPU tests call the product only this way."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from scda_b200 import _lib
    return _lib.load()
