"""CPU checks of the drop-in boundary: the header, the built library and the ctypes binding agree."""
import ctypes
import os
import re

import pytest

from flexam_b200 import lib


def _header_functions():
    src = open(lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fx_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(lib.LIB_PATH):
        lib.build()
    return ctypes.CDLL(lib.LIB_PATH)


def test_header_declares_what_binding_uses():
    assert _header_functions() == sorted(lib.exported_symbols())


def test_library_exports_every_declared_symbol(built):
    for name in _header_functions():
        assert hasattr(built, name), f"{name} declared in include/flexam_b200.h but not exported"


def test_abi_version_and_error_slot(built):
    l = lib.load()
    assert l.fx_abi_version() == 1
    assert isinstance(l.fx_last_error(), bytes)


def test_argument_errors_are_reported_without_a_gpu(built):
    l = lib.load()
    # null pointers are rejected before any CUDA call is made
    st = l.fx_gemm_bf16(None, 8, None, 8, None, None, 8, 1, 8, 8, 0, None, None, 0, None, None)
    assert st == -1 and b"null" in l.fx_last_error()
    st = l.fx_fmha_fwd(None, 0, 0, None, 0, 0, None, 0, 0, None, 0, 0, 1, 1, 1, 1, 1.0, None)
    assert st == -1
    with pytest.raises(lib.FlexamNativeError):
        lib.check(st, "fx_fmha_fwd")


def test_ops_refuse_cpu_tensors():
    import torch
    from flexam_b200 import ops
    a = torch.zeros(8, 8, dtype=torch.bfloat16)
    with pytest.raises(lib.FlexamNativeError):
        ops.gemm(a, a, None, a.clone(), 0)
