import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def pytest_sessionstart(session):
    """A fresh checkout has no built artefacts (they are git-ignored): build them once, exactly as the
    driver's build check does. The PRODUCT never builds or falls back by itself — it fails loudly when
    libmp2p_b200.so is missing (tests/test_abi.py)."""
    so = os.path.join(ROOT, "mp2p_icp_b200", "libmp2p_b200.so")
    if not os.path.exists(so):
        import __graft_entry__

        __graft_entry__.build()


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a GPU would otherwise fail noisily; `-m "not gpu"` never selects them.
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
