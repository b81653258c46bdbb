import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

HAVE_REFERENCE = os.path.isdir(os.environ.get("DGMR_REFERENCE", "/root/reference"))
# the unmodified reference package: /root/reference in the build container, baseline/_ref (pip --target install, travels with gpurun) elsewhere
HAVE_REFERENCE_PKG = HAVE_REFERENCE or os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "dgmr", "dgmr.py"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "refpkg: needs the reference package (/root/reference or baseline/_ref)")


def pytest_collection_modifyitems(config, items):
    have_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not have_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
        if "reference" in item.keywords and not HAVE_REFERENCE:
            item.add_marker(pytest.mark.skip(reason="/root/reference not present"))
        if "refpkg" in item.keywords and not HAVE_REFERENCE_PKG:
            item.add_marker(pytest.mark.skip(reason="reference package not present (baseline/_ref)"))


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


OPTIONS = ("umma_cg", "umma_persist", "umma_persist_r", "patch_pair", "patch_mt", "patch_tg", "prefer_patch", "subpix_wgrad_row", "kwstack", "kwstack_pair", "pairconv", "subpix_rows")


@pytest.fixture
def cuda_backend():
    from skillful_nowcasting_b200 import _lib, ops

    old = _lib.set_backend(None)
    ops.clear_pack_cache()
    be = _lib.backend()
    yield be
    for o in OPTIONS:       # tests may flip launcher options: back to the heuristics
        be.set_option(o, -1)
    ops.config.precision = 0
    ops.config.conv_algo = ops.config.wgrad_algo = 0
    _lib.set_backend(old)
    ops.clear_pack_cache()
