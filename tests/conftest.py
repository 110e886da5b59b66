import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "golden_v1.npz"))


@pytest.fixture(scope="session")
def photo():
    """The reference's fixture image (data/test.png as RGB), uint8 [438, 906, 3]."""
    return np.load(os.path.join(GOLDEN_DIR, "photo_438x906.npz"))["rgb"]


@pytest.fixture(scope="session")
def ref_ext():
    """The unmodified reference extension (oracle/_ref), or skip when the prebuilt .so is absent."""
    from oracle.ref_ext import load_ref
    m = load_ref(build_if_missing=os.path.exists("/root/reference"))
    if m is None:
        pytest.skip("oracle/_ref not built")
    return m


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)
