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


@pytest.fixture(scope="session")
def golden_ops():
    return np.load(os.path.join(GOLDEN, "operators.npz"))


@pytest.fixture(scope="session")
def golden_steps():
    return np.load(os.path.join(GOLDEN, "steps.npz"))


@pytest.fixture(scope="session")
def golden_steps_clip():
    """clip_sample=True cases (tests/golden/make_golden.py::make_steps_clip, from the reference's scheduler files)"""
    return np.load(os.path.join(GOLDEN, "steps_clip.npz"))


def rel_l2(a, b):
    """relative L2 error of a against reference b (numpy or torch)."""
    import torch
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return float(np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0))
