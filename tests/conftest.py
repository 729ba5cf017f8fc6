import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def rtf():
    """The product package with its CUDA library loaded (GPU tests only)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import recommend_tf2_b200 as pkg
    pkg.lib()  # raises if librtf_b200.so is missing: GPU tests never fall back
    return pkg


@pytest.fixture(autouse=True)
def _seed_per_test(request):
    """Every test starts from a generator state derived from its own id: layers that draw their
    initial weights from torch's global generator (Dense, Dice, ...) and unseeded `torch.rand*`
    inputs are then the same in every run, whatever ran before — a tolerance that holds once holds
    always, and a failure is reproducible."""
    import zlib

    import numpy as np
    import torch
    seed = zlib.crc32(request.node.nodeid.encode()) & 0x7FFFFFFF
    torch.manual_seed(seed)
    np.random.seed(seed)
    yield
