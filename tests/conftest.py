import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device: without one they are skipped (with the reason), so that a plain
    `pytest tests` on a CPU-only box is green instead of 200+ cuda-init errors.  On a box WITH a GPU nothing is
    skipped: a missing libdrtk_b200.so must fail loudly there."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="gpu test: no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
