"""The C-ABI library: loads without a GPU, exports every symbol include/*.h declares, validates
arguments on the host.  No kernel is launched here."""
import ctypes
import glob
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        src = open(h).read()
        names |= set(re.findall(r"BC_API\s+[\w\s\*]+?\b(bc_\w+)\s*\(", src))
    return names


def test_header_declares_something():
    assert {"bc_gather", "bc_scatter", "bc_copy_blocks", "bc_transfer", "bc_gather_halo", "bc_gather_halo_tiles",
            "bc_compact_mask", "bc_version", "bc_last_error_string"} <= _declared_symbols()


def test_library_exports_every_declared_symbol():
    from blockcopy import _C

    lib = ctypes.CDLL(_C.LIB_PATH)
    missing = [s for s in sorted(_declared_symbols()) if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    # and the Python binding binds exactly the declared set
    assert set(_C.exported_symbols()) == _declared_symbols()


def test_version_and_build_info():
    from blockcopy import _C

    assert _C.lib().bc_version() == 1
    assert "sm_100a" in _C.build_info()


def test_host_side_argument_errors():
    from blockcopy import _C

    lib = _C.lib()
    dummy = ctypes.c_void_p(0x1000)
    # H not divisible by BS -> BC_ERR_SHAPE (-2), nothing launched
    rc = lib.bc_gather(dummy, dummy, dummy, 1, 1, 8, 30, 64, 32, 0, 0, None)
    assert rc == -2 and b"divisible" in lib.bc_last_error_string()
    # NULL pointer
    assert lib.bc_gather(None, dummy, dummy, 1, 1, 8, 32, 64, 32, 0, 0, None) == -1
    # bad dtype enum
    assert lib.bc_scatter(dummy, dummy, dummy, 1, 1, 8, 32, 64, 32, 7, 0, None) == -3
    # misaligned pointer
    assert lib.bc_gather(ctypes.c_void_p(0x1001), dummy, dummy, 1, 1, 8, 32, 64, 32, 0, 0, None) == -4
    # E == 0 is a no-op (reference: no launch when there is nothing to copy, block_funcs.py:33)
    assert lib.bc_gather(None, None, None, 0, 1, 8, 32, 64, 32, 0, 0, None) == 0
    # aliasing
    assert lib.bc_copy_blocks(dummy, dummy, dummy, dummy, 1, 8, 32, 64, 32, 0, 0, None) == -6
    assert lib.bc_gather_halo_tiles(dummy, dummy, dummy, dummy, dummy, 1, 1, 8, 1, 2, 32, 0, 0, 0, None) == -2


def test_python_shim_raises_reference_exception_types():
    import torch
    from blockcopy import _C

    with pytest.raises(AssertionError):  # CUDA tensors only, no CPU fallback
        _C.gather(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 8, 8), torch.zeros(1, dtype=torch.int32), 1)


def test_missing_library_fails_loudly(monkeypatch):
    from blockcopy import _C

    monkeypatch.setattr(_C, "_lib", None)
    monkeypatch.setattr(_C, "LIB_PATH", "/nonexistent/libblockcopy_sm100.so")
    with pytest.raises(ImportError, match="no CPU"):
        _C.lib()
