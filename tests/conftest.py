import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a host without CUDA (or without the built library) is a clean CPU pass plus skips."""
    try:
        import torch
        ok = torch.cuda.is_available() and os.path.exists(os.path.join(ROOT, "ugaitnet_b200", "libugaitnet_b200.so"))
    except Exception:
        ok = False
    if ok:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device and the built libugaitnet_b200.so")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
