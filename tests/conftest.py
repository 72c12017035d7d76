import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run on the GPU box with -m gpu")


@pytest.fixture(scope="session")
def pkg():
    """The product package (built in-tree)."""
    import __graft_entry__ as ge

    ge.build()
    return importlib.import_module("3dhandposeestimation_b200")


@pytest.fixture(scope="session")
def synth_model(pkg):
    return pkg.assets.synthetic_mano()


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


def load_golden(name):
    import numpy as np

    return dict(np.load(os.path.join(GOLDEN, name)))
