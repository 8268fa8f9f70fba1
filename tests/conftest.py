import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle_mod():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def nufft():
    """The product package (C ABI through ctypes).  GPU tests only."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import nufft_b200
    return nufft_b200
