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


def test_shape_stride_and_alignment_errors_without_a_gpu(built):
    """Empty problems, ragged K / N, short leading dimensions, misaligned pointers, an out-of-table RoPE grid and an
    attention row stride shorter than H*128 are all refused with FX_ERR_ARG and a message before any CUDA call (the
    pointers below are never dereferenced)."""
    l = lib.load()
    P = 0x10000          # any 16-byte aligned non-null value
    cases = [
        (l.fx_gemm_bf16, (P, 64, P, 64, None, P, 64, 0, 64, 64, 0, None, None, 0, None, None), b"empty"),      # M = 0
        (l.fx_gemm_bf16, (P, 64, P, 64, None, P, 64, 16, 60, 64, 0, None, None, 0, None, None), b"multiples of 8"),
        (l.fx_gemm_bf16, (P, 32, P, 64, None, P, 64, 16, 64, 64, 0, None, None, 0, None, None), b"lda"),        # lda < K
        (l.fx_gemm_bf16, (P + 2, 64, P, 64, None, P, 64, 16, 64, 64, 0, None, None, 0, None, None), b"align"),
        (l.fx_fmha_fwd, (P, 0, 256, P, 0, 256, P, 0, 256, P, 0, 256, 1, 2, 0, 128, 1.0, None), b"empty"),       # Lq = 0
        (l.fx_fmha_fwd, (P, 0, 128, P, 0, 256, P, 0, 256, P, 0, 256, 1, 2, 128, 128, 1.0, None), b"stride"),    # < H*128
        (l.fx_fmha_fwd, (P, 0, 260, P, 0, 256, P, 0, 256, P, 0, 256, 1, 2, 128, 128, 1.0, None), b"multiples of 8"),
        (l.fx_ln_affine, (P, P, 4, 100, 1e-6, P, P, None), b"bad shape"),                                        # D % 128
        (l.fx_rmsnorm_rope, (P, 3072, 4, 3072, 1e-6, P, None, P, 2000, 4, 4, 0, 4, None), b"RoPE table"),
        (l.fx_rmsnorm_rope, (P, 1000, 4, 3072, 1e-6, P, None, None, 0, 0, 0, 0, 0, None), b"bad shape"),         # ldx < D
    ]
    for fn, args, needle in cases:
        st = fn(*args)
        msg = l.fx_last_error()
        assert st == -1, (fn.__name__, args, st, msg)
        assert needle in msg, (fn.__name__, msg)
