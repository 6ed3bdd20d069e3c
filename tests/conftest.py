"""Shared fixtures.  `-m "not gpu"` runs on the CPU container (oracle, host logic, C-ABI
surface); `-m gpu` are the parity tests proper and need a B200."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def built():
    """Build every native artefact in-tree once per session (no-op when up to date)."""
    import __graft_entry__ as ge
    ge.build()
    return True


def rel_err(a: np.ndarray, b: np.ndarray) -> float:
    """max |a-b| / max |b| -- the logit metric of the north star ("within 1e-4 rel")."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
