import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

HAVE_REFERENCE = os.path.isdir(os.environ.get("DGMR_REFERENCE", "/root/reference"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")


def pytest_collection_modifyitems(config, items):
    have_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not HAVE_REFERENCE:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))


@pytest.fixture
def emu():
    """Route the package's C-ABI calls to the host emulator (host-logic tests only)."""
    from emu_backend import EmuBackend
    from skillful_nowcasting_b200 import _lib, ops

    old = _lib.set_backend(EmuBackend())
    ops.clear_pack_cache()
    yield
    _lib.set_backend(old)
    ops.clear_pack_cache()


@pytest.fixture
def cuda_backend():
    from skillful_nowcasting_b200 import _lib, ops

    old = _lib.set_backend(None)
    ops.clear_pack_cache()
    yield _lib.backend()
    _lib.set_backend(old)
    ops.clear_pack_cache()
