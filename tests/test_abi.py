"""The C-ABI library loads and exports every symbol include/littlegan_b200.h declares (no GPU)."""
import ctypes
import os
import re

from littlegan_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "littlegan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    if not os.path.exists(_lib.LIB_PATH):
        from littlegan_b200.csrc.build import build
        build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
        assert n in _lib.SIGNATURES, "no ctypes signature for " + n
    assert sorted(_lib.SIGNATURES) == names


def test_version_and_error_channel():
    lib = _lib.load()
    assert lib.lg_abi_version() == 1
    assert isinstance(lib.lg_last_error(), bytes)
    # argument validation happens before any CUDA call, so it works without a GPU
    r = lib.lg_conv2d_fprop(None, None, None, None, None, None, 0, 0, 0, 0, 0, 3, 0, 0, None, None)
    assert r == -1 and b"invalid geometry" in lib.lg_last_error()
    assert lib.lg_gemm(None, None, None, None, 1, 1, 1, 0, 0, 0, 0, 0, None) == -1
    assert lib.lg_pack_conv_weights(None, None, 64, 128, None) == 2 * 25 * 64 * 128 * 2
