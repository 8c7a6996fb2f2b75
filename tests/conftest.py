import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu via gpurun)")


def rel_err(a, b):
    """max|a-b| / max|b| — the parity metric of SURVEY §8d."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-12)).item()


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from clipcap_b200 import _ffi
    _ffi.lib()  # fail loudly if the native library is missing
    return torch.device("cuda:0")
