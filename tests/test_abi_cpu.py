"""The C-ABI library builds for sm_100a without a GPU, loads, and exports exactly the symbols include/b3d.h declares
(no compute calls here)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "segmentation-and-classification-of-brain-tumor-using-3d-unet_b200")


@pytest.fixture(scope="module")
def lib_path():
    import __graft_entry__ as g
    g.build()
    path = os.path.join(PKG, "libb3d.so")
    assert os.path.exists(path)
    return path


def _declared():
    src = open(os.path.join(ROOT, "include", "b3d.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b3d_[a-zA-Z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 35
    for n in names:
        assert hasattr(lib, n), "missing export: " + n


def test_no_undeclared_exports(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (b3d_[a-zA-Z0-9_]+)", out)))
    assert exported == _declared()


def test_error_channel_without_gpu(lib_path):
    lib = ctypes.CDLL(lib_path)
    lib.b3d_last_error_string.restype = ctypes.c_char_p
    assert lib.b3d_version() >= 100
    import torch
    if not torch.cuda.is_available():
        assert lib.b3d_check_device() != 0          # fails loudly, no fallback
        assert len(lib.b3d_last_error_string()) > 0


def test_sass_contains_blackwell_tensor_and_tma_ops(lib_path):
    """tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md table)."""
    out = subprocess.run(["cuobjdump", "-sass", lib_path], capture_output=True, text=True).stdout
    assert re.search(r"UTC\w*MMA", out), "no tcgen05.mma in SASS"
    assert "LDTM" in out and "UTMALDG" in out
    assert not re.search(r"(?<!UTC)HMMA", out), "legacy mma.sync path must not be present"


def test_product_never_imports_oracle():
    for f in os.listdir(PKG):
        if f.endswith(".py"):
            src = open(os.path.join(PKG, f)).read()
            assert "oracle" not in src, f
