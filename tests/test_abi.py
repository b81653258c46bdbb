"""The C-ABI shared library loads and exports every symbol include/dgmr_b200.h declares (no GPU needed);
the product path refuses CPU tensors instead of falling back."""
import ctypes
import os

import pytest
import torch

from skillful_nowcasting_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_build_and_symbols():
    import __graft_entry__ as ge

    ge.build()
    assert os.path.exists(_lib.LIB_PATH)
    decls = _lib.parse_header()
    assert len(decls) >= 35
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in decls:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    lib.dgmr_abi_version.restype = ctypes.c_int
    assert lib.dgmr_abi_version() == 1


def test_header_cites_reference_for_every_family():
    src = open(_lib.HEADER).read()
    for needle in ("ConvGRU.py", "common.py", "generators.py", "discriminators.py", "losses.py", "parametrizations.py", "Attention.py"):
        assert needle in src


def test_no_cpu_fallback():
    """Without a CUDA tensor the product path raises; it never computes on the CPU."""
    old = _lib.set_backend(None)
    try:
        be = _lib.backend()  # CudaBackend: dlopen works without a GPU
        x = torch.zeros(8)
        with pytest.raises(RuntimeError, match="CUDA tensor"):
            be.relu_fwd(x, torch.empty(8))
    finally:
        _lib.set_backend(old)


def test_product_package_does_not_import_oracle():
    pkg = os.path.join(ROOT, "skillful_nowcasting_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("# oracle", ""), f"{f} mentions the oracle"
                assert "emu_backend" not in text
