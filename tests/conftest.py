import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def has_gpu():
    return _has_gpu()


def pytest_runtest_setup(item):
    """A gpu-marked test on a box without a CUDA device must fail loudly rather than pass or skip (the product has no
    CPU fallback to test): `pytest -m gpu` on a CPU-only box is an error report, `pytest -m "not gpu"` is the CPU suite."""
    if item.get_closest_marker("gpu") and not _has_gpu():
        pytest.fail("this test needs a CUDA device and none is visible (torch.cuda.is_available() is False); "
                    "run `pytest -m 'not gpu'` on a CPU-only box", pytrace=False)
