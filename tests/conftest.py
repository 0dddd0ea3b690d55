import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (test infrastructure): oracle/sodso_oracle.cpp via ctypes."""
    from oracle import oracle as O

    O.lib()
    return O


@pytest.fixture(scope="session")
def real_scans():
    return dict(np.load(os.path.join(GOLDEN, "real_scans_seq06.npz")))


@pytest.fixture(scope="session")
def gpu_ctx():
    import torch

    if not torch.cuda.is_available():
        pytest.fail("this test needs a GPU: the product path has no CPU fallback")
    from so_dso_place_recognition_b200 import api

    return api.default_context(0)
