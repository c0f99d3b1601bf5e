import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "fake-video-detection-engine_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def _has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(params=["smem", "mma"])
def block_stage(request):
    """Runs a GPU test once per build of the fused kernel's block stage (include/v5ela.h v5ela_set_block_stage): the calling
    thread's cached handle is switched for the duration of the test."""
    import torch
    from v5ela.batch import get_handle

    hd = get_handle(torch.cuda.current_device())
    old = hd.block_stage
    hd.block_stage = request.param
    yield request.param
    hd.block_stage = old
