import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


class Golden(object):
    """npz fixture with 'case/name' keys."""

    def __init__(self, fname):
        self.z = np.load(os.path.join(GOLDEN, fname))

    def case(self, tag):
        pre = tag + "/"
        return {k[len(pre):]: self.z[k] for k in self.z.files if k.startswith(pre)}

    def __getitem__(self, k):
        return self.z[k]

    @property
    def files(self):
        return self.z.files


@pytest.fixture(scope="session")
def golden():
    return Golden


@pytest.fixture(autouse=True)
def _fp32_library_math():
    """Parity is judged against fp32 references: every test starts with TF32 off for cuDNN and cuBLAS (torch's default
    allows it for convolutions), independent of test order.  The TF32 kernels (decoder tail, phase convolutions) are
    selected by ``torch.backends.cudnn.allow_tf32``; their tests switch it on explicitly and this restores the state."""
    import torch
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
