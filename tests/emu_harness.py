"""Loads the SIMT-emulated kernel build and lets the CPU tests drive the product's Python layer with it.

TEST INFRASTRUCTURE: `emulated()` patches, for the duration of a test, the device check of the binding layer
so that host tensors reach `tests/emu/_build/libmdn_loss_emu.so`.  The product never does this.
"""
import contextlib
import os
import sys
from unittest import mock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))

_lib = None


def emu_library():
    global _lib
    if _lib is None:
        import build_emu
        from mdn_sfm_b200 import _cabi
        _lib = _cabi.Library(build_emu.build())
    return _lib


@contextlib.contextmanager
def emulated():
    from mdn_sfm_b200 import _cabi

    def check_tensor(t, dtype=None, what="tensor"):
        import torch
        dtype = dtype or torch.float32
        assert t.dtype == dtype, (what, t.dtype)
        return t

    with mock.patch.object(_cabi, "_lib", emu_library()), mock.patch.object(_cabi, "check_tensor", check_tensor):
        yield emu_library()
