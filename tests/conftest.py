import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def cuda_device():
    """GPU tests fail (not skip) when the native path cannot run: a silent
    fallback would void the parity claim."""
    import torch

    assert torch.cuda.is_available(), "gpu-marked test but torch sees no CUDA device"
    from tristage_rag_b200 import _lib

    assert _lib.lib().ts_device_count() >= 1, "libtristage sees no sm_100 device"
    return 0
