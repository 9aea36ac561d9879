import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this environment")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    """Loader for the committed golden vectors (tests/golden/*.npz)."""
    import numpy as np

    def load(name: str):
        return np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    return load


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the native libraries once per session (no-op when up to date)."""
    import __graft_entry__ as g
    g.build()
