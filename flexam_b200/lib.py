"""ctypes binding of the C-ABI library declared in ``include/flexam_b200.h``.

The library is the product: there is no Python/torch fallback. Importing this module never touches the
GPU; the first call into an ``fx_*`` entry point does. A missing ``libflexam_b200.so`` raises at load time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC_DIR = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(CSRC_DIR, "libflexam_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "flexam_b200.h")

_vp, _i, _i64, _f = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes; every function returns int status except the two noted below.
SIGNATURES = {
    "fx_abi_version": [],
    "fx_check_device": [_i],
    "fx_gemm_bf16": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _i64, _vp, _vp],
    "fx_ln_modulate": [_vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _vp],
    "fx_ln_affine": [_vp, _vp, _i, _i, _f, _vp, _vp, _vp],
    "fx_modulation_tables": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "fx_ln_scale_shift": [_vp, _vp, _i, _i, _f, _vp, _vp, _i64, _vp, _vp],
    "fx_rmsnorm_rope": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fx_fmha_fwd": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _f, _vp],
    "fx_qkv_norm_rope_scatter": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, C.POINTER(_vp), _i, _i, _i64,
                                 _i, _vp],
    "fx_fmha_fwd_scatter": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, C.POINTER(_vp), _i, _i, _i64, _i64, _i,
                            _i, _i, _i, _f, _vp],
    "fx_patchify": [C.POINTER(_vp), C.POINTER(_i), C.POINTER(_i), _i, _i, _i, _i, _vp, _i64, _vp],
    "fx_unpatchify": [_vp, _i64, _vp, _i, _i, _i, _i, _vp],
    "fx_sinusoid": [_vp, _vp, _i, _i, _vp],
    "fx_linear_f32": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _vp],
    "fx_linear_f32_tc": [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i, _i, _i, _i, _i, _vp, _vp],
    "fx_dedup_f32": [_vp, _i, _i, _vp, _vp, _vp, _vp],
    "fx_nchw_to_nhwc": [_vp, _vp, _i64, _i, _i, _i64, _vp],
    "fx_im2col3x3": [_vp, _vp, _i, _i, _i, _i, _vp],
    "fx_groupnorm_silu": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "fx_groupnorm_partials": [_vp, _i, _i64, _i, _i, _vp, _vp],
    "fx_groupnorm_silu_partials": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _i64, _vp, _vp, _vp, _i, _i, _i64, _vp, _vp],
    "fx_conv_gemm_bf16": [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fx_nchw_to_nhwc_padded": [_vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "fx_cfg_euler_step": [_vp, _vp, _f, _f, _vp, _vp, _vp, _i64, _vp],
    "fx_swap01_bf16": [_vp, _i64, _vp, _i, _i, _i, _vp],
    "fx_add_f32": [_vp, _vp, _i64, _vp],
    "fx_sub_f32": [_vp, _vp, _vp, _i64, _vp],
    "fx_fingerprint": [_vp, _vp, _i, _i, _vp, _vp],
    "fx_tune": [C.c_char_p, _i],
    # umT5 text encoder (flexam_b200/text_encoder.py)
    "fx_embedding_bf16": [_vp, _vp, _vp, _i64, _i, _i64, _vp],
    "fx_t5_layernorm": [_vp, _vp, _vp, _i, _i, _f, _vp],
    "fx_t5_attention": [_vp, _i64, _vp, _vp, _vp, _i64, _i, _i, _i, _vp],
    "fx_add_bf16": [_vp, _vp, _i64, _vp],
    "fx_gated_gelu_bf16": [_vp, _i64, _vp, _i64, _vp, _i64, _i64, _i, _vp],
    # Wan2.2 VAE decoder (flexam_b200/vae.py)
    "fx_vae_norm_act": [_vp, _i64, _vp, _vp, _i64, _i64, _i, _i, _i, _i, _i, _i, _vp],
    "fx_vae_upsample2x": [_vp, _vp, _i, _i, _i, _i, _vp],
    "fx_vae_time_interleave": [_vp, _vp, _i, _i64, _i, _vp],
    "fx_vae_dupup_add": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fx_vae_halo_push": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fx_softmax_rows_f32": [_vp, _i64, _vp, _i64, _i, _i, _f, _vp],
    "fx_vae_unpatchify": [_vp, _i64, _vp, _i, _i, _i, _i, _i, _vp],
    "fx_vae_patchify": [_vp, _vp, _i64, _i, _i, _i, _i, _i, _vp],
    "fx_vae_avgdown_add": [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "fx_cast_f32_to_bf16": [_vp, _vp, _i64, _vp],
    "fx_cast_bf16_to_f32": [_vp, _vp, _i64, _vp],
    # fp32 verification mode (flexam_b200/precise.py)
    "fx_split3_f32": [_vp, _i64, _i, _i, _vp, _vp],
    "fx_join3_f32": [_vp, _i64, _vp, _vp],
    "fx_ln_f32": [_vp, _vp, _i, _i, _f, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _i, _vp, _vp, _vp],
    "fx_rmsnorm_rope_f32": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp],
    "fx_gelu_f32": [_vp, _i64, _vp],
    "fx_gated_residual_f32": [_vp, _vp, _i, _i, _vp, _vp, _i64, _vp, _vp],
    "fx_attention_f32": [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _f, _vp],
    "fx_groupnorm_silu_f32": [_vp, _i64, _i, _i, _f, _vp, _vp, _vp, _vp, _vp, _vp],
}

FX_EPI_BF16, FX_EPI_GELU_BF16, FX_EPI_F32, FX_EPI_RESID_F32, FX_EPI_F32_EXACT = 0, 1, 2, 3, 4

_lib = None


class FlexamNativeError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the library in-tree with nvcc for sm_100a (works without a GPU)."""
    cmd = ["make", "-C", CSRC_DIR, "-j", str(os.cpu_count() or 4)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout)
        print(r.stderr)
    if r.returncode != 0:
        raise FlexamNativeError("building libflexam_b200.so failed")
    return LIB_PATH


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FlexamNativeError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(there is no CPU/torch fallback for the native path)")
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = C.c_int
    lib.fx_last_error.argtypes = []
    lib.fx_last_error.restype = C.c_char_p
    ver = lib.fx_abi_version()
    if ver != 1:
        raise FlexamNativeError(f"ABI version mismatch: library {ver}, binding 1")
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != 0:
        msg = load().fx_last_error().decode("utf-8", "replace")
        raise FlexamNativeError(f"{what} failed with status {status}: {msg}")


def exported_symbols() -> list[str]:
    return list(SIGNATURES.keys()) + ["fx_last_error"]
