"""CPU: the C-ABI library builds, loads, and exports every symbol include/cgs_b200.h declares;
argument validation returns error codes without touching a GPU; the product refuses CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    from cgs_b200 import _lib
    return _lib.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cgs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cgs_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/cgs_b200.h but not exported"


def test_binding_table_matches_header(lib):
    from cgs_b200 import _lib
    assert set(_lib.EXPORTS) | {"cgs_last_error", "cgs_version"} == set(declared_symbols())
    assert lib.cgs_version() >= 100


def test_struct_layout_matches_header():
    from cgs_b200 import _lib
    # cgs_src: 4 x int32 + 3 pointers; conv args embed it first
    assert ctypes.sizeof(_lib.Src) == 16 + 3 * 8
    assert _lib.Conv3x3Args.src.offset == 0
    assert ctypes.sizeof(_lib.Conv3x3Args) % 8 == 0


def test_validation_errors_without_gpu(lib):
    from cgs_b200 import _lib
    a = _lib.Conv3x3Args()
    assert lib.cgs_conv3x3(ctypes.byref(a), None) == -1          # CGS_EINVAL: null operand
    assert b"conv3x3" in lib.cgs_last_error()
    assert lib.cgs_dense_fwd(None, None, None, 1, 1, 1, None, None) == -1
    assert lib.cgs_conv3x3(None, None) == -1


def test_no_cpu_fallback():
    from cgs_b200._lib import CgsError
    from cgs_b200.nets import NewCritic, UnetDecoder
    c, m = NewCritic(), UnetDecoder()
    x = torch.rand(2, 3, 64, 64)
    with pytest.raises(CgsError):
        c(x)
    with pytest.raises(CgsError):
        m(x, [torch.zeros(2, 8, 32, 32)] * 5)
    with pytest.raises(ValueError):
        c(torch.rand(2, 3, 32, 32))
