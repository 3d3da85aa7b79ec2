import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _seed():
    # same global fixture as the reference's tests/conftest.py:6-10 (grad disabled by default)
    np.random.seed(42)
    torch.manual_seed(42)
    prev = torch.is_grad_enabled()
    torch.set_grad_enabled(False)
    yield
    torch.set_grad_enabled(prev)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def rel_err(a, b):
    """max |a-b| / max(|b|, 1): relative where the value is large, absolute near zero."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    if a.numel() == 0:
        return 0.0
    both_inf = torch.isinf(a) & torch.isinf(b) & (a.sign() == b.sign())
    d = torch.where(both_inf, torch.zeros_like(a), (a - b).abs())
    return float((d / b.abs().clamp_min(1.0)).max())


def norm_err(a, b):
    """||a-b||_inf / ||b||_inf -- for gradient tensors whose entries span many magnitudes."""
    a = torch.as_tensor(a, dtype=torch.float64).cpu()
    b = torch.as_tensor(b, dtype=torch.float64).cpu()
    if a.numel() == 0:
        return 0.0
    pinned = ~torch.isnan(b)     # the reference yields NaN gradients for NaN inputs: those are unpinned
    if not bool(pinned.any()):
        return 0.0
    a, b = a[pinned], b[pinned]
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
