"""Host-side operator layer: torch tensors in, C-ABI calls out.

PyTorch is used for device memory and streams only; every function below validates its arguments and
forwards raw pointers + the current CUDA stream to ``libflexam_b200.so`` (``include/flexam_b200.h``).
Nothing here computes on the host or falls back to torch ops.
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional, Sequence

import torch

from . import lib as _l
from .lib import FX_EPI_BF16, FX_EPI_F32, FX_EPI_F32_EXACT, FX_EPI_GELU_BF16, FX_EPI_RESID_F32  # noqa: F401

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32


_scoped_stream: Optional[int] = None


def _stream() -> int:
    return _scoped_stream if _scoped_stream is not None else torch.cuda.current_stream().cuda_stream


@contextlib.contextmanager
def stream_scope():
    """Resolve torch's current CUDA stream once for a whole forward (``torch.cuda.current_stream()`` costs ~14 us, a
    third of the host time of a 500-launch step). Nothing inside the native forward switches streams."""
    global _scoped_stream
    prev = _scoped_stream
    if torch.cuda.is_available():
        _scoped_stream = torch.cuda.current_stream().cuda_stream
    try:
        yield
    finally:
        _scoped_stream = prev


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _req(t: torch.Tensor, dtype, name: str, inner_contig: bool = True) -> None:
    if not t.is_cuda:
        raise _l.FlexamNativeError(f"{name}: expected a CUDA tensor (the native path has no CPU fallback)")
    if t.dtype != dtype:
        raise _l.FlexamNativeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if inner_contig and t.dim() > 0 and t.stride(-1) != 1:
        raise _l.FlexamNativeError(f"{name}: innermost dimension must be contiguous")


def gemm(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, epilogue: int,
         gate_mod: Optional[torch.Tensor] = None, gate_e: Optional[torch.Tensor] = None,
         row_idx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = epilogue(a[M,K] @ w[N,K]^T + bias). See fx_gemm_bf16."""
    _req(a, bf16, "gemm.a"), _req(w, bf16, "gemm.w")
    M, K = a.shape
    N = w.shape[0]
    if w.shape[1] != K or out.shape[0] != M or out.shape[1] != N:
        raise _l.FlexamNativeError(f"gemm: shape mismatch a{tuple(a.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    _req(out, bf16 if epilogue in (FX_EPI_BF16, FX_EPI_GELU_BF16) else f32, "gemm.out")
    if bias is not None:
        _req(bias, bf16, "gemm.bias")
    ge_stride = 0
    if gate_mod is not None:
        _req(gate_mod, f32, "gemm.gate_mod")
    if gate_e is not None:
        _req(gate_e, f32, "gemm.gate_e")
        ge_stride = gate_e.stride(0) if gate_e.dim() == 2 else 0
    if row_idx is not None:
        _req(row_idx, i32, "gemm.row_idx")
    st = _l.load().fx_gemm_bf16(_p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), M, N, K,
                                epilogue, _p(gate_mod), _p(gate_e), ge_stride, _p(row_idx), _stream())
    _l.check(st, "fx_gemm_bf16")
    return out


def ln_modulate(x: torch.Tensor, out: torch.Tensor, eps: float, shift_mod: torch.Tensor, scale_mod: torch.Tensor,
                shift_e: torch.Tensor, scale_e: torch.Tensor, e_stride: int, row_idx: Optional[torch.Tensor],
                dens_mod: Optional[torch.Tensor], dens: Optional[torch.Tensor], dens_stride: int,
                rows_per_batch: int) -> torch.Tensor:
    _req(x, f32, "ln_modulate.x"), _req(out, bf16, "ln_modulate.out")
    for n, t in (("shift_mod", shift_mod), ("scale_mod", scale_mod), ("shift_e", shift_e), ("scale_e", scale_e)):
        _req(t, f32, "ln_modulate." + n)
    M, D = x.shape
    st = _l.load().fx_ln_modulate(_p(x), _p(out), M, D, eps, _p(shift_mod), _p(scale_mod), _p(shift_e), _p(scale_e),
                                  e_stride, _p(row_idx), _p(dens_mod), _p(dens), dens_stride, rows_per_batch,
                                  _stream())
    _l.check(st, "fx_ln_modulate")
    return out


def modulation_tables(mod: torch.Tensor, dmod: torch.Tensor, e0: torch.Tensor, de0: torch.Tensor,
                      tab: torch.Tensor) -> torch.Tensor:
    """tab f32 [2, U*B, 2, D] from mod [6,D], dmod [2,D], e0 [U,6,D], de0 [B,2,D] (all f32, contiguous)."""
    for n, t in (("mod", mod), ("dmod", dmod), ("e0", e0), ("de0", de0), ("tab", tab)):
        _req(t, f32, "modulation_tables." + n)
        if not t.is_contiguous():
            raise _l.FlexamNativeError(f"modulation_tables.{n}: contiguous tensor required")
    U, B, D = e0.shape[0], de0.shape[0], mod.shape[1]
    if tuple(tab.shape) != (2, U * B, 2, D) or tuple(e0.shape) != (U, 6, D) or tuple(de0.shape) != (B, 2, D):
        raise _l.FlexamNativeError("modulation_tables: shape mismatch")
    _l.check(_l.load().fx_modulation_tables(_p(mod), _p(dmod), _p(e0), _p(de0), U, B, D, _p(tab), _stream()),
             "fx_modulation_tables")
    return tab


def ln_scale_shift(x: torch.Tensor, out: torch.Tensor, eps: float, scale: torch.Tensor, shift: torch.Tensor,
                   row_stride: int, row_idx: Optional[torch.Tensor]) -> torch.Tensor:
    _req(x, f32, "ln_scale_shift.x"), _req(out, bf16, "ln_scale_shift.out")
    _req(scale, f32, "ln_scale_shift.scale"), _req(shift, f32, "ln_scale_shift.shift")
    if row_idx is not None:
        _req(row_idx, i32, "ln_scale_shift.row_idx")
    M, D = x.shape
    st = _l.load().fx_ln_scale_shift(_p(x), _p(out), M, D, eps, _p(scale), _p(shift), row_stride, _p(row_idx), _stream())
    _l.check(st, "fx_ln_scale_shift")
    return out


def ln_affine(x: torch.Tensor, out: torch.Tensor, eps: float, gamma: torch.Tensor, beta: torch.Tensor) -> torch.Tensor:
    _req(x, f32, "ln_affine.x"), _req(out, bf16, "ln_affine.out")
    _req(gamma, bf16, "ln_affine.gamma"), _req(beta, bf16, "ln_affine.beta")
    M, D = x.shape
    _l.check(_l.load().fx_ln_affine(_p(x), _p(out), M, D, eps, _p(gamma), _p(beta), _stream()), "fx_ln_affine")
    return out


def rmsnorm_rope(x: torch.Tensor, weight: torch.Tensor, eps: float, freqs: Optional[torch.Tensor] = None,
                 grid: Sequence[int] = (0, 0, 0), tok_offset: int = 0, rows_per_batch: int = 0,
                 weight2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """In place on x: bf16 [M, D] view (row stride free). With weight2, x is an [M, 2D] view whose second D columns
    are normed with weight2 in the same launch (q and k of the packed projection output)."""
    _req(x, bf16, "rmsnorm_rope.x"), _req(weight, bf16, "rmsnorm_rope.weight")
    if freqs is not None:
        _req(freqs, f32, "rmsnorm_rope.freqs")
    M, D = x.shape
    if weight2 is not None:
        _req(weight2, bf16, "rmsnorm_rope.weight2")
        D //= 2
    st = _l.load().fx_rmsnorm_rope(_p(x), x.stride(0), M, D, eps, _p(weight), _p(weight2), _p(freqs), int(grid[0]),
                                   int(grid[1]), int(grid[2]), tok_offset, rows_per_batch if rows_per_batch > 0 else M,
                                   _stream())
    _l.check(st, "fx_rmsnorm_rope")
    return x


def fmha(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, scale: float) -> torch.Tensor:
    """q/out: [B, Lq, H, 128] views, k/v: [B, Lk, H, 128] views (bf16, head dim and heads contiguous)."""
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _req(t, bf16, "fmha." + n)
        if t.dim() != 4 or t.shape[3] != 128 or t.stride(2) != 128:
            raise _l.FlexamNativeError(f"fmha.{n}: expected [B, L, H, 128] with contiguous heads, got "
                                       f"{tuple(t.shape)} strides {t.stride()}")
    B, Lq, H, _ = q.shape
    Lk = k.shape[1]
    st = _l.load().fx_fmha_fwd(_p(q), q.stride(0), q.stride(1), _p(k), k.stride(0), k.stride(1), _p(v), v.stride(0),
                               v.stride(1), _p(out), out.stride(0), out.stride(1), B, H, Lq, Lk, scale, _stream())
    _l.check(st, "fx_fmha_fwd")
    return out


def qkv_norm_rope_scatter(qkv: torch.Tensor, D: int, weight_q: torch.Tensor, weight_k: torch.Tensor, eps: float,
                          freqs: torch.Tensor, grid: Sequence[int], tok_offset: int, rows_per_batch: int,
                          peer_ptrs: Sequence[int], heads_per_peer: int, dst_rows: int, dst_row0: int) -> None:
    """qkv: bf16 [M, 3D] packed projection output of this rank's token slice. q and k are RMS-normed + rotated, v is
    moved as is, and every head lands in the exchange buffer of the rank that owns it (peer memory). See
    fx_qkv_norm_rope_scatter."""
    _req(qkv, bf16, "qkv_norm_rope_scatter.qkv"), _req(weight_q, bf16, "qkv_norm_rope_scatter.weight_q")
    _req(weight_k, bf16, "qkv_norm_rope_scatter.weight_k"), _req(freqs, f32, "qkv_norm_rope_scatter.freqs")
    n = len(peer_ptrs)
    ptrs = (C.c_void_p * n)(*peer_ptrs)
    st = _l.load().fx_qkv_norm_rope_scatter(_p(qkv), qkv.stride(0), qkv.shape[0], D, eps, _p(weight_q), _p(weight_k),
                                            _p(freqs), int(grid[0]), int(grid[1]), int(grid[2]), tok_offset,
                                            rows_per_batch, ptrs, n, heads_per_peer, dst_rows, dst_row0, _stream())
    _l.check(st, "fx_qkv_norm_rope_scatter")


def fmha_scatter(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, peer_ptrs: Sequence[int], rows_per_peer: int,
                 o_stride_b: int, o_stride_l: int, scale: float) -> None:
    """fmha whose output row i goes to rank i // rows_per_peer (peer_ptrs: device addresses of every rank's [B, rows,
    H_total, 128] attention-output buffer, already offset to this rank's first head). See fx_fmha_fwd_scatter."""
    for n, t in (("q", q), ("k", k), ("v", v)):
        _req(t, bf16, "fmha_scatter." + n)
        if t.dim() != 4 or t.shape[3] != 128 or t.stride(2) != 128:
            raise _l.FlexamNativeError(f"fmha_scatter.{n}: expected [B, L, H, 128] with contiguous heads, got "
                                       f"{tuple(t.shape)} strides {t.stride()}")
    B, Lq, H, _ = q.shape
    Lk = k.shape[1]
    n = len(peer_ptrs)
    ptrs = (C.c_void_p * n)(*peer_ptrs)
    st = _l.load().fx_fmha_fwd_scatter(_p(q), q.stride(0), q.stride(1), _p(k), k.stride(0), k.stride(1), _p(v),
                                       v.stride(0), v.stride(1), ptrs, n, rows_per_peer, o_stride_b, o_stride_l, B, H,
                                       Lq, Lk, scale, _stream())
    _l.check(st, "fx_fmha_fwd_scatter")


def patchify(srcs: Sequence[torch.Tensor], chan_last: Sequence[bool], F: int, H: int, W: int,
             rows: torch.Tensor) -> torch.Tensor:
    n = len(srcs)
    ptrs = (C.c_void_p * n)(*[s.data_ptr() for s in srcs])
    chans, cl = [], []
    for s, last in zip(srcs, chan_last):
        _req(s, bf16, "patchify.src", inner_contig=False)
        if not s.is_contiguous():
            raise _l.FlexamNativeError("patchify: sources must be contiguous")
        chans.append(s.shape[-1] if last else s.shape[0])
        cl.append(1 if last else 0)
    _req(rows, bf16, "patchify.rows")
    st = _l.load().fx_patchify(ptrs, (C.c_int * n)(*chans), (C.c_int * n)(*cl), n, F, H, W, _p(rows), rows.stride(0),
                               _stream())
    _l.check(st, "fx_patchify")
    return rows


def unpatchify(head: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """head: bf16 [tokens, >=4C] view; out: bf16 [C, F, H, W] contiguous."""
    _req(head, bf16, "unpatchify.head"), _req(out, bf16, "unpatchify.out")
    Cc, F, H, W = out.shape
    _l.check(_l.load().fx_unpatchify(_p(head), head.stride(0), _p(out), Cc, F, H, W, _stream()), "fx_unpatchify")
    return out


def sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    _req(t, f32, "sinusoid.t")
    out = torch.empty((t.numel(), dim), dtype=f32, device=t.device)
    _l.check(_l.load().fx_sinusoid(_p(t), _p(out), t.numel(), dim, _stream()), "fx_sinusoid")
    return out


def linear_f32(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act_in: int = 0) -> torch.Tensor:
    _req(x, f32, "linear_f32.x"), _req(w, bf16, "linear_f32.w")
    if bias is not None:
        _req(bias, bf16, "linear_f32.bias")
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty((M, N), dtype=f32, device=x.device)
    st = _l.load().fx_linear_f32(_p(x), x.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), M, N, K,
                                 act_in, _stream())
    _l.check(st, "fx_linear_f32")
    return out


def linear_f32_tc(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act_in: int = 0, planes: int = 2,
                  out: Optional[torch.Tensor] = None, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 linear for many rows on the tensor cores (see fx_linear_f32_tc). ws: bf16 workspace >= M*planes*K."""
    _req(x, f32, "linear_f32_tc.x"), _req(w, bf16, "linear_f32_tc.w")
    if bias is not None:
        _req(bias, bf16, "linear_f32_tc.bias")
    M, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((M, N), dtype=f32, device=x.device)
    _req(out, f32, "linear_f32_tc.out")
    if ws is None:
        ws = torch.empty((M, planes * K), dtype=bf16, device=x.device)
    _req(ws, bf16, "linear_f32_tc.ws")
    if ws.numel() < M * planes * K or not ws.is_contiguous() or out.shape != (M, N):
        raise _l.FlexamNativeError("linear_f32_tc: workspace too small / not contiguous, or out shape mismatch")
    st = _l.load().fx_linear_f32_tc(_p(x), x.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), M, N, K,
                                    act_in, planes, _p(ws), _stream())
    _l.check(st, "fx_linear_f32_tc")
    return out


def dedup_f32(t: torch.Tensor, cap: int, uniq: torch.Tensor, inv: torch.Tensor, count: torch.Tensor) -> None:
    """Distinct values of the flat fp32 tensor t (first-appearance order, at most cap) without a host sync."""
    _req(t, f32, "dedup_f32.t"), _req(uniq, f32, "dedup_f32.uniq"), _req(inv, i32, "dedup_f32.inv")
    _req(count, i32, "dedup_f32.count")
    if not t.is_contiguous() or uniq.numel() < cap or inv.numel() < t.numel():
        raise _l.FlexamNativeError("dedup_f32: contiguous t and uniq[cap], inv[n] required")
    _l.check(_l.load().fx_dedup_f32(_p(t), t.numel(), cap, _p(uniq), _p(inv), _p(count), _stream()), "fx_dedup_f32")


class FingerprintTable:
    """Device arrays (addresses, byte sizes) of a fixed list of tensors; keeps the tensors alive."""

    def __init__(self, tensors):
        self.tensors = list(tensors)
        for t in self.tensors:
            if not t.is_cuda or not t.is_contiguous() or t.data_ptr() % 16 != 0:
                raise _l.FlexamNativeError("fingerprint_table: contiguous, 16-byte aligned CUDA tensors required")
        dev = self.tensors[0].device
        self.ptrs = torch.tensor([t.data_ptr() for t in self.tensors], dtype=torch.int64, device=dev)
        self.nbytes = torch.tensor([t.numel() * t.element_size() for t in self.tensors], dtype=torch.int64, device=dev)


def fingerprint_table(tensors) -> Optional[FingerprintTable]:
    """None for CPU tensors: an engine may be built before ``module.to('cuda')`` (the reference pipelines do); nothing
    native can run on them, and the table is rebuilt when the parameters arrive on the device."""
    tensors = list(tensors)
    if not tensors or not tensors[0].is_cuda:
        return None
    return FingerprintTable(tensors)


def fingerprint(table: FingerprintTable, stride: int) -> torch.Tensor:
    """int64 [n]: sampled hash-sum of every tensor of the table (see fx_fingerprint); a fresh tensor per call."""
    n = table.ptrs.numel()
    out = torch.empty((n,), dtype=torch.int64, device=table.ptrs.device)
    _l.check(_l.load().fx_fingerprint(_p(table.ptrs), _p(table.nbytes), n, stride, _p(out), _stream()), "fx_fingerprint")
    return out


def tune(name: str, value: int) -> None:
    """Developer knob by name (see fx_tune)."""
    _l.check(_l.load().fx_tune(name.encode(), int(value)), "fx_tune")


def nchw_to_nhwc(src: torch.Tensor, dst: torch.Tensor, c0: int) -> torch.Tensor:
    """src: bf16 [C, P] contiguous view; dst: bf16 [P, ld] receives columns [c0, c0+C)."""
    _req(src, bf16, "nchw_to_nhwc.src"), _req(dst, bf16, "nchw_to_nhwc.dst")
    Cc, P = src.shape
    _l.check(_l.load().fx_nchw_to_nhwc(_p(src), _p(dst), dst.stride(0), c0, Cc, P, _stream()), "fx_nchw_to_nhwc")
    return dst


def im2col3x3(x: torch.Tensor, F: int, H: int, W: int, rows: torch.Tensor) -> torch.Tensor:
    _req(x, bf16, "im2col3x3.x"), _req(rows, bf16, "im2col3x3.rows")
    Cc = x.shape[-1]
    _l.check(_l.load().fx_im2col3x3(_p(x), _p(rows), F, H, W, Cc, _stream()), "fx_im2col3x3")
    return rows


def groupnorm_silu(x: torch.Tensor, groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
                   resid: Optional[torch.Tensor], y_f32: Optional[torch.Tensor], y_bf16: Optional[torch.Tensor],
                   stats: torch.Tensor) -> None:
    _req(x, bf16, "groupnorm_silu.x")
    P, Cc = x.shape
    st = _l.load().fx_groupnorm_silu(_p(x), P, Cc, groups, eps, _p(gamma), _p(beta), _p(resid), _p(y_f32), _p(y_bf16),
                                     _p(stats), _stream())
    _l.check(st, "fx_groupnorm_silu")


def groupnorm_partials(x: torch.Tensor, frames: int, groups: int, partials: torch.Tensor) -> torch.Tensor:
    """partials f64 [frames, groups, 2] = per-(frame, group) sum / sum of squares of x bf16 [frames*pp, C]."""
    _req(x, bf16, "groupnorm_partials.x"), _req(partials, torch.float64, "groupnorm_partials.partials")
    P, Cc = x.shape
    if P % frames != 0 or tuple(partials.shape) != (frames, groups, 2) or not partials.is_contiguous():
        raise _l.FlexamNativeError("groupnorm_partials: x rows must split into frames; partials [frames, groups, 2]")
    st = _l.load().fx_groupnorm_partials(_p(x), frames, P // frames, Cc, groups, _p(partials), _stream())
    _l.check(st, "fx_groupnorm_partials")
    return partials


def conv_gemm(act: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor, T: int, H: int,
              W: int, kt: int, ks: int, epilogue: int = FX_EPI_BF16, stride_s: int = 1, stride_t: int = 1) -> torch.Tensor:
    """Implicit-GEMM convolution (see fx_conv_gemm_bf16). act: bf16 zero-padded channel-last [(T+kt-1)*Hp*Wp, Cin]
    contiguous; w: bf16 [Cout, kt*ks*ks*Cin]; out: dense [T*H*W, Cout] view (row stride free)."""
    _req(act, bf16, "conv_gemm.act"), _req(w, bf16, "conv_gemm.w")
    Cin, Cout = act.shape[1], w.shape[0]
    Hp, Wp = (H + 2, W + 2) if ks == 3 else (H, W)
    if not act.is_contiguous() or act.shape[0] != (T + kt - 1) * Hp * Wp or w.shape[1] != kt * ks * ks * Cin \
            or not w.is_contiguous() or tuple(out.shape) != ((T // stride_t) * (H // stride_s) * (W // stride_s), Cout):
        raise _l.FlexamNativeError(f"conv_gemm: shape mismatch act{tuple(act.shape)} w{tuple(w.shape)} out{tuple(out.shape)}")
    _req(out, bf16 if epilogue in (FX_EPI_BF16, FX_EPI_GELU_BF16) else f32, "conv_gemm.out")
    if bias is not None:
        _req(bias, bf16, "conv_gemm.bias")
    st = _l.load().fx_conv_gemm_bf16(_p(act), _p(w), _p(bias), _p(out), out.stride(0), T, H, W, Cin, Cout, kt, ks,
                                     stride_s, stride_t, epilogue, _stream())
    _l.check(st, "fx_conv_gemm_bf16")
    return out


def nchw_to_nhwc_padded(src: torch.Tensor, dst: torch.Tensor, c0: int, F: int, H: int, W: int) -> torch.Tensor:
    """src: bf16 [C, F*H*W] contiguous; dst: bf16 [F*(H+2)*(W+2), ld] zero-padded channel-last grid, columns [c0, c0+C)."""
    _req(src, bf16, "nchw_to_nhwc_padded.src"), _req(dst, bf16, "nchw_to_nhwc_padded.dst")
    Cc, P = src.shape
    if P != F * H * W or dst.shape[0] != F * (H + 2) * (W + 2) or not src.is_contiguous():
        raise _l.FlexamNativeError("nchw_to_nhwc_padded: shape mismatch")
    _l.check(_l.load().fx_nchw_to_nhwc_padded(_p(src), _p(dst), dst.stride(0), c0, Cc, F, H, W, _stream()),
             "fx_nchw_to_nhwc_padded")
    return dst


def groupnorm_silu_partials(x: torch.Tensor, groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
                            partials: torch.Tensor, pix_per_frame: int, resid: Optional[torch.Tensor],
                            y_f32: Optional[torch.Tensor], y_bf16: Optional[torch.Tensor], stats: torch.Tensor,
                            pad_hw: Optional[Sequence[int]] = None) -> None:
    """GroupNorm + SiLU (+ resid) of the local pixels x [P, C] with the statistics of ALL frames: partials f64
    [Ft, groups, 2] in frame order (see fx_groupnorm_silu_partials). pad_hw=(H, W): y_bf16 is the zero-padded grid
    [frames*(H+2)*(W+2), ld] the next implicit-GEMM convolution reads (interior positions are written)."""
    _req(x, bf16, "groupnorm_silu_partials.x"), _req(partials, torch.float64, "groupnorm_silu_partials.partials")
    if partials.dim() != 3 or partials.shape[1] != groups or not partials.is_contiguous():
        raise _l.FlexamNativeError("groupnorm_silu_partials: partials must be contiguous [Ft, groups, 2]")
    P, Cc = x.shape
    ph, pw = (int(pad_hw[0]), int(pad_hw[1])) if pad_hw is not None else (0, 0)
    ld_b = y_bf16.stride(0) if (y_bf16 is not None and pad_hw is not None) else 0
    st = _l.load().fx_groupnorm_silu_partials(_p(x), P, Cc, groups, eps, _p(gamma), _p(beta), _p(partials),
                                              partials.shape[0], pix_per_frame, _p(resid), _p(y_f32), _p(y_bf16), ph, pw,
                                              ld_b, _p(stats), _stream())
    _l.check(st, "fx_groupnorm_silu_partials")


def cfg_euler_step(vu: torch.Tensor, vc: torch.Tensor, guidance: float, dsigma: float, lat: torch.Tensor,
                   mask: Optional[torch.Tensor], pinned: Optional[torch.Tensor]) -> torch.Tensor:
    _req(vu, bf16, "cfg_euler_step.vu"), _req(vc, bf16, "cfg_euler_step.vc"), _req(lat, f32, "cfg_euler_step.lat")
    st = _l.load().fx_cfg_euler_step(_p(vu), _p(vc), guidance, dsigma, _p(lat), _p(mask), _p(pinned), lat.numel(),
                                     _stream())
    _l.check(st, "fx_cfg_euler_step")
    return lat


def swap01(src: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out[b, a, :] = src[a, b, :]; src: bf16 [A, B, inner] (row stride of A free), out: contiguous [B, A, inner]."""
    _req(src, bf16, "swap01.src"), _req(out, bf16, "swap01.out")
    A, B, inner = src.shape
    if src.stride(1) != inner or not out.is_contiguous() or out.shape != (B, A, inner):
        raise _l.FlexamNativeError(f"swap01: bad layout src{tuple(src.shape)}/{src.stride()} out{tuple(out.shape)}")
    _l.check(_l.load().fx_swap01_bf16(_p(src), src.stride(0), _p(out), A, B, inner, _stream()), "fx_swap01_bf16")
    return out


def add_(dst: torch.Tensor, src: torch.Tensor) -> torch.Tensor:
    _req(dst, f32, "add_.dst"), _req(src, f32, "add_.src")
    if not (dst.is_contiguous() and src.is_contiguous() and dst.numel() == src.numel()):
        raise _l.FlexamNativeError("add_: contiguous tensors of equal size required")
    _l.check(_l.load().fx_add_f32(_p(dst), _p(src), dst.numel(), _stream()), "fx_add_f32")
    return dst


def sub(a: torch.Tensor, b: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    for n, t in (("a", a), ("b", b), ("out", out)):
        _req(t, f32, "sub." + n)
        if not t.is_contiguous() or t.numel() != a.numel():
            raise _l.FlexamNativeError("sub: contiguous tensors of equal size required")
    _l.check(_l.load().fx_sub_f32(_p(out), _p(a), _p(b), a.numel(), _stream()), "fx_sub_f32")
    return out


# ----------------------------------------------------------------------------------------------------------
# fp32 verification mode (see include/flexam_b200.h and flexam_b200/precise.py)
# ----------------------------------------------------------------------------------------------------------
def split3(x: torch.Tensor, planes: torch.Tensor) -> torch.Tensor:
    """planes bf16 [3, M, K] (hi, mid, lo) with hi + mid + lo == x exactly; x f32 [M, K] (row stride free)."""
    _req(x, f32, "split3.x"), _req(planes, bf16, "split3.planes")
    M, K = x.shape
    if tuple(planes.shape) != (3, M, K) or not planes.is_contiguous():
        raise _l.FlexamNativeError(f"split3: planes must be contiguous [3, {M}, {K}], got {tuple(planes.shape)}")
    _l.check(_l.load().fx_split3_f32(_p(x), x.stride(0), M, K, _p(planes), _stream()), "fx_split3_f32")
    return planes


def join3(planes: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _req(planes, bf16, "join3.planes"), _req(out, f32, "join3.out")
    if planes.shape[0] != 3 or planes[0].numel() != out.numel() or not planes.is_contiguous() or not out.is_contiguous():
        raise _l.FlexamNativeError("join3: planes must be contiguous [3, n...] matching a contiguous out")
    _l.check(_l.load().fx_join3_f32(_p(planes), out.numel(), _p(out), _stream()), "fx_join3_f32")
    return out


def ln_f32(x: torch.Tensor, out: torch.Tensor, eps: float, shift_mod=None, scale_mod=None, shift_e=None, scale_e=None,
           e_stride: int = 0, row_idx=None, dens_mod=None, dens=None, dens_stride: int = 0, rows_per_batch: int = 0,
           gamma=None, beta=None) -> torch.Tensor:
    _req(x, f32, "ln_f32.x"), _req(out, f32, "ln_f32.out")
    M, D = x.shape
    st = _l.load().fx_ln_f32(_p(x), _p(out), M, D, eps, _p(shift_mod), _p(scale_mod), _p(shift_e), _p(scale_e), e_stride,
                             _p(row_idx), _p(dens_mod), _p(dens), dens_stride, rows_per_batch, _p(gamma), _p(beta),
                             _stream())
    _l.check(st, "fx_ln_f32")
    return out


def rmsnorm_rope_f32(x: torch.Tensor, weight: torch.Tensor, eps: float, freqs: Optional[torch.Tensor] = None,
                     grid: Sequence[int] = (0, 0, 0), tok_offset: int = 0, rows_per_batch: int = 0,
                     weight2: Optional[torch.Tensor] = None) -> torch.Tensor:
    _req(x, f32, "rmsnorm_rope_f32.x"), _req(weight, bf16, "rmsnorm_rope_f32.weight")
    M, D = x.shape
    if weight2 is not None:
        D //= 2
    st = _l.load().fx_rmsnorm_rope_f32(_p(x), x.stride(0), M, D, eps, _p(weight), _p(weight2), _p(freqs), int(grid[0]),
                                       int(grid[1]), int(grid[2]), tok_offset,
                                       rows_per_batch if rows_per_batch > 0 else M, _stream())
    _l.check(st, "fx_rmsnorm_rope_f32")
    return x


def gelu_f32_(x: torch.Tensor) -> torch.Tensor:
    _req(x, f32, "gelu_f32.x")
    if not x.is_contiguous():
        raise _l.FlexamNativeError("gelu_f32: contiguous tensor required")
    _l.check(_l.load().fx_gelu_f32(_p(x), x.numel(), _stream()), "fx_gelu_f32")
    return x


def gated_residual_f32_(x: torch.Tensor, y: torch.Tensor, gate_mod=None, gate_e=None, row_idx=None) -> torch.Tensor:
    _req(x, f32, "gated_residual_f32.x"), _req(y, f32, "gated_residual_f32.y")
    if x.shape != y.shape or not x.is_contiguous() or not y.is_contiguous():
        raise _l.FlexamNativeError("gated_residual_f32: x and y must be contiguous and equal-shaped")
    M, N = x.shape
    ge_stride = gate_e.stride(0) if (gate_e is not None and gate_e.dim() == 2) else 0
    st = _l.load().fx_gated_residual_f32(_p(x), _p(y), M, N, _p(gate_mod), _p(gate_e), ge_stride, _p(row_idx), _stream())
    _l.check(st, "fx_gated_residual_f32")
    return x


def attention_f32(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, scale: float) -> torch.Tensor:
    for n, t in (("q", q), ("k", k), ("v", v), ("out", out)):
        _req(t, f32, "attention_f32." + n)
        if t.dim() != 4 or t.shape[3] != 128 or t.stride(2) != 128:
            raise _l.FlexamNativeError(f"attention_f32.{n}: expected [B, L, H, 128] with contiguous heads")
    B, Lq, H, _ = q.shape
    st = _l.load().fx_attention_f32(_p(q), q.stride(0), q.stride(1), _p(k), k.stride(0), k.stride(1), _p(v),
                                    v.stride(0), v.stride(1), _p(out), out.stride(0), out.stride(1), B, H, Lq,
                                    k.shape[1], scale, _stream())
    _l.check(st, "fx_attention_f32")
    return out


def groupnorm_silu_f32(x: torch.Tensor, groups: int, eps: float, gamma: torch.Tensor, beta: torch.Tensor,
                       resid: Optional[torch.Tensor], y: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
    _req(x, f32, "groupnorm_silu_f32.x"), _req(y, f32, "groupnorm_silu_f32.y")
    _req(gamma, bf16, "groupnorm_silu_f32.gamma"), _req(beta, bf16, "groupnorm_silu_f32.beta")
    P, Cc = x.shape
    st = _l.load().fx_groupnorm_silu_f32(_p(x), P, Cc, groups, eps, _p(gamma), _p(beta), _p(resid), _p(y), _p(stats),
                                         _stream())
    _l.check(st, "fx_groupnorm_silu_f32")
    return y


# ----------------------------------------------------------------------------------------------------------
# umT5 text encoder operators (see include/flexam_b200.h and flexam_b200/text_encoder.py)
# ----------------------------------------------------------------------------------------------------------
def embedding(ids: torch.Tensor, table: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    _req(table, bf16, "embedding.table"), _req(out, bf16, "embedding.out")
    if ids.dtype != torch.int64 or not ids.is_cuda or not ids.is_contiguous() or not out.is_contiguous():
        raise _l.FlexamNativeError("embedding: contiguous int64 CUDA ids and a contiguous out required")
    rows, D = ids.numel(), table.shape[1]
    _l.check(_l.load().fx_embedding_bf16(_p(ids), _p(table), _p(out), rows, D, table.shape[0], _stream()),
             "fx_embedding_bf16")
    return out


def t5_layernorm(x: torch.Tensor, weight: torch.Tensor, out: torch.Tensor, eps: float = 1e-6) -> torch.Tensor:
    _req(x, bf16, "t5_layernorm.x"), _req(weight, bf16, "t5_layernorm.weight"), _req(out, bf16, "t5_layernorm.out")
    if not (x.is_contiguous() and out.is_contiguous()):
        raise _l.FlexamNativeError("t5_layernorm: contiguous tensors required")
    M, D = x.shape
    _l.check(_l.load().fx_t5_layernorm(_p(x), _p(weight), _p(out), M, D, eps, _stream()), "fx_t5_layernorm")
    return out


def t5_attention(qkv: torch.Tensor, bias_rel: torch.Tensor, mask: Optional[torch.Tensor], out: torch.Tensor, B: int,
                 L: int, H: int) -> torch.Tensor:
    _req(qkv, bf16, "t5_attention.qkv"), _req(bias_rel, bf16, "t5_attention.bias_rel"), _req(out, bf16, "t5_attention.out")
    if mask is not None:
        _req(mask, i32, "t5_attention.mask")
    if tuple(bias_rel.shape) != (H, 2 * L - 1) or not bias_rel.is_contiguous() or qkv.shape[0] != B * L:
        raise _l.FlexamNativeError("t5_attention: bias_rel must be contiguous [H, 2L-1]; qkv rows must be B*L")
    st = _l.load().fx_t5_attention(_p(qkv), qkv.stride(0), _p(bias_rel), _p(mask), _p(out), out.stride(0), B, L, H,
                                   _stream())
    _l.check(st, "fx_t5_attention")
    return out


def add_bf16_(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    _req(x, bf16, "add_bf16.x"), _req(y, bf16, "add_bf16.y")
    if not (x.is_contiguous() and y.is_contiguous() and x.numel() == y.numel()):
        raise _l.FlexamNativeError("add_bf16: contiguous tensors of equal size required")
    _l.check(_l.load().fx_add_bf16(_p(x), _p(y), x.numel(), _stream()), "fx_add_bf16")
    return x


def gated_gelu(fc1: torch.Tensor, gate: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
    """out = bf16(fc1 * GELU(gate)) on [M, N] views (row strides free: the two halves of one packed GEMM output)."""
    for n, t in (("fc1", fc1), ("gate", gate), ("out", out)):
        _req(t, bf16, "gated_gelu." + n)
        if t.dim() != 2 or t.shape != fc1.shape:
            raise _l.FlexamNativeError("gated_gelu: 2-D tensors of equal shape required")
    M, N = fc1.shape
    st = _l.load().fx_gated_gelu_bf16(_p(fc1), fc1.stride(0), _p(gate), gate.stride(0), _p(out), out.stride(0), M, N,
                                      _stream())
    _l.check(st, "fx_gated_gelu_bf16")
    return out


# ----------------------------------------------------------------------------------------------------------
# Wan2.2 VAE decoder operators (see include/flexam_b200.h and flexam_b200/vae.py)
# ----------------------------------------------------------------------------------------------------------
def vae_norm_act(x: torch.Tensor, gamma: Optional[torch.Tensor], out: torch.Tensor, H: int, W: int, pad: int,
                 frame0: int, silu: bool) -> torch.Tensor:
    """x: bf16 [npix, C] (row stride free); out: bf16 [(frames)*(H+2pad)*(W+2pad), ldo] grid (or dense rows, pad 0)."""
    _req(x, bf16, "vae_norm_act.x"), _req(out, bf16, "vae_norm_act.out")
    if gamma is not None:
        _req(gamma, bf16, "vae_norm_act.gamma")
    npix, Cc = x.shape
    frames = npix // (H * W)
    if out.shape[0] < (frame0 + frames) * (H + 2 * pad) * (W + 2 * pad) or out.shape[1] < Cc:
        raise _l.FlexamNativeError("vae_norm_act: output grid too small")
    st = _l.load().fx_vae_norm_act(_p(x), x.stride(0), _p(gamma), _p(out), out.stride(0), npix, Cc, H, W, pad, frame0,
                                   1 if silu else 0, _stream())
    _l.check(st, "fx_vae_norm_act")
    return out


def vae_upsample2x(x: torch.Tensor, out: torch.Tensor, Fr: int, H: int, W: int) -> torch.Tensor:
    _req(x, bf16, "vae_upsample2x.x"), _req(out, bf16, "vae_upsample2x.out")
    Cc = x.shape[1]
    if not (x.is_contiguous() and out.is_contiguous()) or x.shape[0] != Fr * H * W or \
            tuple(out.shape) != (Fr * (2 * H + 2) * (2 * W + 2), Cc):
        raise _l.FlexamNativeError("vae_upsample2x: shape mismatch")
    _l.check(_l.load().fx_vae_upsample2x(_p(x), _p(out), Fr, H, W, Cc, _stream()), "fx_vae_upsample2x")
    return out


def vae_time_interleave(y: torch.Tensor, x: torch.Tensor, T: int, P: int) -> torch.Tensor:
    _req(y, bf16, "vae_time_interleave.y"), _req(x, bf16, "vae_time_interleave.x")
    Cc = x.shape[1]
    if not (y.is_contiguous() and x.is_contiguous()) or tuple(y.shape) != (T * P, 2 * Cc) or x.shape[0] != 2 * T * P:
        raise _l.FlexamNativeError("vae_time_interleave: shape mismatch")
    _l.check(_l.load().fx_vae_time_interleave(_p(y), _p(x), T, P, Cc, _stream()), "fx_vae_time_interleave")
    return x


def vae_dupup_add_(main: torch.Tensor, x: torch.Tensor, Tout: int, H: int, W: int, ft: int, drop: int) -> torch.Tensor:
    """main: bf16 [Tout*2H*2W, Cout] += DupUp3D(x: bf16 [T*H*W, Cin]) (see fx_vae_dupup_add)."""
    _req(main, bf16, "vae_dupup_add.main"), _req(x, bf16, "vae_dupup_add.x")
    if not (main.is_contiguous() and x.is_contiguous()) or main.shape[0] != Tout * 4 * H * W:
        raise _l.FlexamNativeError("vae_dupup_add: shape mismatch")
    st = _l.load().fx_vae_dupup_add(_p(main), _p(x), Tout, H, W, x.shape[1], main.shape[1], ft, drop, _stream())
    _l.check(st, "fx_vae_dupup_add")
    return main


def vae_halo_push(grid: torch.Tensor, up_ptr: int, dn_ptr: int, frame0: int, T: int, Hp: int, Wp: int) -> None:
    """grid: bf16 [>= (frame0+T)*Hp*Wp, C] contiguous; up_ptr / dn_ptr: device addresses of the SAME grid on the rank
    above / below (0 = none). See fx_vae_halo_push."""
    _req(grid, bf16, "vae_halo_push.grid")
    if not grid.is_contiguous() or grid.shape[0] < (frame0 + T) * Hp * Wp:
        raise _l.FlexamNativeError("vae_halo_push: grid too small")
    st = _l.load().fx_vae_halo_push(_p(grid), C.c_void_p(up_ptr or None), C.c_void_p(dn_ptr or None), frame0, T,
                                    Hp, Wp, grid.shape[1], _stream())
    _l.check(st, "fx_vae_halo_push")


def softmax_rows(s: torch.Tensor, p: torch.Tensor, scale: float) -> torch.Tensor:
    _req(s, f32, "softmax_rows.s"), _req(p, bf16, "softmax_rows.p")
    rows, cols = s.shape
    if tuple(p.shape) != (rows, cols):
        raise _l.FlexamNativeError("softmax_rows: shape mismatch")
    _l.check(_l.load().fx_softmax_rows_f32(_p(s), s.stride(0), _p(p), p.stride(0), rows, cols, scale, _stream()),
             "fx_softmax_rows_f32")
    return p


def vae_unpatchify(y: torch.Tensor, video: torch.Tensor, T: int, H: int, W: int, frame0: int) -> torch.Tensor:
    """y: bf16 [T*H*W, >=12]; video: bf16 [3, Ttot, 2H, 2W] contiguous, frames [frame0, frame0+T) written, clamped."""
    _req(y, bf16, "vae_unpatchify.y"), _req(video, bf16, "vae_unpatchify.video")
    if not video.is_contiguous() or video.shape[0] != 3 or tuple(video.shape[2:]) != (2 * H, 2 * W) or y.shape[0] != T * H * W:
        raise _l.FlexamNativeError("vae_unpatchify: shape mismatch")
    st = _l.load().fx_vae_unpatchify(_p(y), y.stride(0), _p(video), T, H, W, video.shape[1], frame0, _stream())
    _l.check(st, "fx_vae_unpatchify")
    return video


def vae_patchify(video: torch.Tensor, rows: torch.Tensor, T: int, h: int, w: int, frame0: int) -> torch.Tensor:
    """video: bf16 [3, Ttot, 2h, 2w] contiguous; rows: bf16 [T*h*w, >=12] (columns [0, 12) written)."""
    _req(video, bf16, "vae_patchify.video"), _req(rows, bf16, "vae_patchify.rows")
    if not video.is_contiguous() or video.shape[0] != 3 or tuple(video.shape[2:]) != (2 * h, 2 * w) or rows.shape[0] != T * h * w:
        raise _l.FlexamNativeError("vae_patchify: shape mismatch")
    st = _l.load().fx_vae_patchify(_p(video), _p(rows), rows.stride(0), T, h, w, video.shape[1], frame0, _stream())
    _l.check(st, "fx_vae_patchify")
    return rows


def vae_avgdown_add_(main: torch.Tensor, x: torch.Tensor, T: int, H: int, W: int, ft: int, fs: int) -> torch.Tensor:
    """main: bf16 [ceil(T/ft)*(H/fs)*(W/fs), Cout] += AvgDown3D(x: bf16 [T*H*W, Cin]) (see fx_vae_avgdown_add)."""
    _req(main, bf16, "vae_avgdown_add.main"), _req(x, bf16, "vae_avgdown_add.x")
    To = -(-T // ft)
    if not (main.is_contiguous() and x.is_contiguous()) or main.shape[0] != To * (H // fs) * (W // fs) or x.shape[0] != T * H * W:
        raise _l.FlexamNativeError("vae_avgdown_add: shape mismatch")
    st = _l.load().fx_vae_avgdown_add(_p(main), _p(x), T, H, W, x.shape[1], main.shape[1], ft, fs, _stream())
    _l.check(st, "fx_vae_avgdown_add")
    return main
