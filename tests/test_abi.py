"""The C-ABI library: loads without a GPU, exports every symbol include/*.h declares, and validates
arguments before touching CUDA (no compute calls here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    from jarvis_hybridnet_b200 import _lib
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "jarvis_hybridnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(jhn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    from jarvis_hybridnet_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"


def test_abi_version(lib):
    from jarvis_hybridnet_b200 import _lib
    assert lib.jhn_abi_version() == _lib.ABI_VERSION == 6


def test_shape_validation_without_gpu(lib):
    n = ctypes.c_size_t()
    assert lib.jhn_reproject_workspace_bytes(1, 12, 23, 130, 72, 0, ctypes.byref(n)) == 0 and n.value > 0
    assert lib.jhn_reproject_workspace_bytes(1, 12, 23, 130, 70, 0, ctypes.byref(n)) == -2      # G % 4 != 0
    assert b"multiple of 4" in lib.jhn_last_error()
    assert lib.jhn_reproject_workspace_bytes(1, 12, 25, 130, 72, 0, ctypes.byref(n)) == -2      # K > 24
    assert lib.jhn_reproject_workspace_bytes(0, 12, 23, 130, 72, 0, ctypes.byref(n)) == -2      # B < 1
    assert lib.jhn_reproject_workspace_bytes(1, 12, 23, 130, 72, 0, None) == -1


def test_null_and_count_validation(lib):
    out = ctypes.c_void_p()
    assert lib.jhn_v2v_create(None, 24, 23, 0, None, ctypes.byref(out)) == -1
    arr = (ctypes.c_void_p * 3)(1, 2, 3)
    assert lib.jhn_v2v_create(arr, 3, 23, 0, None, ctypes.byref(out)) == -2
    assert b"24" in lib.jhn_last_error()
    assert lib.jhn_centroid_reduce(None, 1, 1, 4, 2.0, 16.0, None, None, None, None, None) == -1
    assert lib.jhn_v2v_forward(None, None, 0, 1, 8, None, None, 0, None) == -1
    lib.jhn_v2v_destroy(None)          # must be a no-op


def test_workspace_grows_with_batch(lib):
    a, b = ctypes.c_size_t(), ctypes.c_size_t()
    lib.jhn_reproject_workspace_bytes(1, 12, 23, 130, 72, 1, ctypes.byref(a))
    lib.jhn_reproject_workspace_bytes(4, 12, 23, 130, 72, 1, ctypes.byref(b))
    assert 3.9 * a.value < b.value <= 4 * a.value + 4096


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from jarvis_hybridnet_b200 import _lib
    assert _lib.load().jhn_check_device(0) == -4
    assert b"CUDA error" in _lib.load().jhn_last_error()
