import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def rel_l2(a, b):
    import torch
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    d = (a - b).norm()
    n = b.norm()
    return float(d / n) if n > 0 else float(d)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name)))
    return load


@pytest.fixture(autouse=True)
def _fp32_parity(request):
    """Every GPU parity test compares with fp32 results of the reference / oracle on the host: keep cuDNN and cuBLAS
    out of TF32 for all of them, so that no test depends on an earlier module having switched it off."""
    if request.node.get_closest_marker("gpu") is not None:
        import torch
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
    yield
