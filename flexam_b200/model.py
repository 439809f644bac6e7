"""Host-side mirror of the reference transformer interface, backed only by the C-ABI kernels.

``Wan2_2Transformer3DModel_FlexAM`` here has the reference's constructor kwargs, parameter tree (state_dict keys
and shapes), ``forward(x, t, context, seq_len, clip_fea, y, y_camera, full_ref, subject_ref,
cond_flag, additional_control, density)`` signature and feature toggles
(FlexAM/models/wan_transformer3d_FlexAM.py:526-1123, :1335-1438), but its forward is ``NativeEngine.forward``: a
fixed sequence of ``fx_*`` launches on the caller's stream. ``install(module)`` rebinds the forward of an existing
reference module instance the same way the reference's own ``enable_multi_gpus_inference`` does (:801-815).

There is no torch/cuDNN/flash-attn compute on this path and no CPU fallback: torch provides device buffers,
views and streams. Parameter packing (q|k|v concatenation, conv-weight flattening, fp32 copies of the modulation
rows) happens once per weight version and is re-done if a parameter is edited in place (LoRA merge).
"""
from __future__ import annotations

import contextlib
import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from . import teacache as _tc
from .lib import FX_EPI_BF16, FX_EPI_F32, FX_EPI_GELU_BF16, FX_EPI_RESID_F32, FlexamNativeError

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32


# ----------------------------------------------------------------------------------------------------------
# parameter tree (names / shapes of the reference state_dict)
# ----------------------------------------------------------------------------------------------------------
def param_shapes(cfg: dict) -> Dict[str, tuple]:
    D, Fd = cfg["dim"], cfg["ffn_dim"]
    pt, ph, pw = cfg["patch_size"]
    out: Dict[str, tuple] = {}

    def lin(p, o, i):
        out[p + ".weight"] = (o, i)
        out[p + ".bias"] = (o,)

    out["patch_embedding.weight"] = (D, cfg["in_dim"], pt, ph, pw)
    out["patch_embedding.bias"] = (D,)
    lin("text_embedding.0", D, cfg["text_dim"]), lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg["freq_dim"]), lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    lin("density_embedding.0", D, cfg["freq_dim"]), lin("density_embedding.2", D, D)
    lin("density_projection.1", 2 * D, D)
    for i in range(cfg["num_layers"]):
        b = f"blocks.{i}."
        out[b + "modulation"] = (1, 6, D)
        out[b + "modulation_density"] = (1, 2, D)
        for att in ("self_attn", "cross_attn"):
            for nm in "qkvo":
                lin(b + f"{att}.{nm}", D, D)
            out[b + f"{att}.norm_q.weight"] = (D,)
            out[b + f"{att}.norm_k.weight"] = (D,)
        out[b + "norm3.weight"] = (D,)
        out[b + "norm3.bias"] = (D,)
        lin(b + "ffn.0", Fd, D), lin(b + "ffn.2", D, Fd)
    lin("head.head", cfg["out_dim"] * pt * ph * pw, D)
    out["head.modulation"] = (1, 2, D)
    out["head.modulation_density"] = (1, 1, D)
    out["ref_conv.weight"] = (D, cfg["in_dim_ref_conv"], ph, pw)
    out["ref_conv.bias"] = (D,)
    for j, (ci, co) in enumerate([(cfg["in_dim_cnn_block"], 192), (192, 192), (192, 96), (96, 96)], start=1):
        out[f"cnn_conv{j}.0.weight"] = (co, ci, 1, 3, 3)
        out[f"cnn_conv{j}.0.bias"] = (co,)
        out[f"cnn_conv{j}.1.weight"] = (co,)
        out[f"cnn_conv{j}.1.bias"] = (co,)
    out["cnn_conv5.weight"] = (cfg["out_dim_cnn_block"], 96, 1, 1, 1)
    out["cnn_conv5.bias"] = (cfg["out_dim_cnn_block"],)
    return out


def rope_table(head_dim: int, theta: float = 10000.0, max_len: int = 1024, riflex: Optional[dict] = None) -> torch.Tensor:
    """fp32 [max_len, head_dim/2, 2] (cos, sin) table the RMSNorm+RoPE kernels read (rope_table_f64 rounded once)."""
    return rope_table_f64(head_dim, theta, max_len, riflex).to(f32).contiguous()


def rope_table_f64(head_dim: int, theta: float = 10000.0, max_len: int = 1024, riflex: Optional[dict] = None) -> torch.Tensor:
    """float64 [max_len, head_dim/2, 2] (cos, sin), the construction of rope_params (:44-52, :655-665);
    RIFLEx (:56-113, :774-788) changes one temporal frequency."""
    d = head_dim
    cols = []
    for axis, dim in enumerate((d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6))):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
        if axis == 0 and riflex is not None:
            k = riflex["k"]
            inv[k - 1] = 0.9 * 2 * math.pi / riflex["L_test"]
            if riflex.get("L_test_scale") is not None:
                inv[k - 1] = inv[k - 1] / riflex["L_test_scale"]
        cols.append(torch.outer(torch.arange(max_len, dtype=torch.float64), inv))
    ang = torch.cat(cols, dim=1)
    return torch.stack([ang.cos(), ang.sin()], dim=-1)


def rope_complex(head_dim: int, riflex: Optional[dict] = None) -> torch.Tensor:
    """The reference's ``freqs`` attribute: complex128 [1024, head_dim/2] (:655-665; enable_riflex :774-788)."""
    t = rope_table_f64(head_dim, riflex=riflex)
    return torch.complex(t[..., 0], t[..., 1])


# ----------------------------------------------------------------------------------------------------------
# the engine: packed weights + workspaces + the launch sequence
# ----------------------------------------------------------------------------------------------------------
class NativeEngine:
    """Runs the denoising step for a parameter mapping with the reference's state_dict keys (bf16, on one GPU)."""

    def __init__(self, params: Dict[str, torch.Tensor], cfg: dict, device: torch.device):
        self.cfg = dict(cfg)
        self.device = torch.device(device)
        self.params = params
        self.D = cfg["dim"]
        self.H = cfg["num_heads"]
        if self.D // self.H != 128:
            raise FlexamNativeError(f"native path supports head_dim 128 only (dim {self.D}, heads {self.H})")
        self.eps = float(cfg["eps"])
        self.freqs = rope_table(128).to(self.device)
        self.par = None                # flexam_b200.dist.Parallel: CFG-branch x Ulysses sequence parallelism
        self.cache_static = True       # hoist step-invariant work (context, cross K/V, CNN fuser) across calls
        self._static_key = None
        self._static = {}
        self._ws = {}
        self._versions = None
        self.launches = 0              # kernels launched by the last forward (bench.py reports it)
        self.timing = None             # list -> per-launch CUDA events for the GEMM / attention families
        # Per-call checks that need ONE device->host read (weight fingerprint, content of the step-invariant inputs,
        # number of distinct timesteps). A host that vouches for all three (DenoiseLoop after its first step) sets
        # `trusted` and passes `t_dedup`: the forward then enqueues without any host synchronisation.
        self.trusted = False
        self.max_table_timesteps = 64  # more distinct per-token timesteps than this: per-token modulation path
        self.mlp_planes = 2            # bf16 planes of the fp32 input in the tensor-core time MLP (per-token path)
        self.host_reads = 0            # device->host reads of the last forward (tests / bench)
        self._pack()

    # -- weights -------------------------------------------------------------------------------------------
    def _pvers(self):
        return tuple(p._version for p in self.params.values())

    def _pack(self):
        P, D, dev = self.params, self.D, self.device
        for k, v in P.items():
            if v.device != dev or v.dtype != bf16:
                raise FlexamNativeError(f"parameter {k}: expected bf16 on {dev}, got {v.dtype} on {v.device}")
        L = self.cfg["num_layers"]
        self.blk = []
        for i in range(L):
            b = f"blocks.{i}."
            sa, ca = b + "self_attn.", b + "cross_attn."
            w = {
                "wqkv": torch.cat([P[sa + "q.weight"], P[sa + "k.weight"], P[sa + "v.weight"]], 0).contiguous(),
                "bqkv": torch.cat([P[sa + "q.bias"], P[sa + "k.bias"], P[sa + "v.bias"]], 0).contiguous(),
                "wo": P[sa + "o.weight"], "bo": P[sa + "o.bias"],
                "nq": P[sa + "norm_q.weight"], "nk": P[sa + "norm_k.weight"],
                "n3w": P[b + "norm3.weight"], "n3b": P[b + "norm3.bias"],
                "cwq": P[ca + "q.weight"], "cbq": P[ca + "q.bias"],
                "cwkv": torch.cat([P[ca + "k.weight"], P[ca + "v.weight"]], 0).contiguous(),
                "cbkv": torch.cat([P[ca + "k.bias"], P[ca + "v.bias"]], 0).contiguous(),
                "cwo": P[ca + "o.weight"], "cbo": P[ca + "o.bias"],
                "cnq": P[ca + "norm_q.weight"], "cnk": P[ca + "norm_k.weight"],
                "w1": P[b + "ffn.0.weight"], "b1": P[b + "ffn.0.bias"],
                "w2": P[b + "ffn.2.weight"], "b2": P[b + "ffn.2.bias"],
                "mod": P[b + "modulation"][0].to(f32).contiguous(),            # [6, D]
                "dmod": P[b + "modulation_density"][0].to(f32).contiguous(),   # [2, D]
            }
            self.blk.append(w)
        self.w_patch = P["patch_embedding.weight"].flatten(1).contiguous()     # [D, in_dim*4], K order (c,q,r)
        self.w_ref = P["ref_conv.weight"].flatten(1).contiguous()
        self.w_cnn = [P[f"cnn_conv{j}.0.weight"].flatten(1).contiguous() for j in range(1, 5)]  # [Co, Ci*9] (c,kh,kw)
        self.w_cnn5 = P["cnn_conv5.weight"].flatten(1).contiguous()
        # the same 3x3 weights for the implicit-GEMM convolution: K order (kh, kw, c) with c zero-padded to a multiple of 64
        self.w_cnn_taps = []
        for j in range(1, 5):
            w = P[f"cnn_conv{j}.0.weight"][:, :, 0]                       # [Co, Ci, 3, 3]
            co, ci = w.shape[:2]
            cp = -(-ci // 64) * 64
            wt = torch.zeros((co, 3, 3, cp), dtype=bf16, device=dev)
            wt[..., :ci] = w.permute(0, 2, 3, 1)
            self.w_cnn_taps.append(wt.view(co, 9 * cp).contiguous())
        self.head_mod = P["head.modulation"][0].to(f32).contiguous()           # [2, D]
        self.head_dmod = P["head.modulation_density"][0, 0].to(f32).contiguous()
        self._versions = self._pvers()
        self._static_key = None
        # sampled fingerprint of every parameter as packed: `param.data += ...` edits (merge_lora) bump no version counter
        self._fp_table = ops.fingerprint_table(list(P.values()))
        self._fp_ref = ops.fingerprint(self._fp_table, self.FP_STRIDE) if self._fp_table is not None else None

    FP_STRIDE = 256  # every 256th 16-byte word: ~40 MB read for the 5 B-parameter model, a dense edit hits every sample

    def refresh_if_modified(self):
        if self._pvers() != self._versions:
            self._pack()

    def refresh(self, params: Optional[Dict[str, torch.Tensor]] = None):
        """Re-derive everything that is a COPY of a weight (packed q|k|v, cross k|v, fp32 modulation rows) and drop the
        step-invariant cache; with ``params`` also adopt new parameter storages. Version counters catch in-place torch
        ops on the parameters, but not edits through ``param.data`` (the reference's merge_lora / unmerge_lora,
        FlexAM/utils/lora_utils.py:481-485, :595-599): hosts call ``flexam_b200.model.refresh(module)`` after those."""
        if params is not None:
            self.params = params
            dev = next(iter(params.values())).device
            if dev != self.device:      # engine created before module.to(device): follow the parameters
                self.device = dev
                self.freqs = self.freqs.to(dev)
                self._ws.clear()
                self._static, self._static_key = {}, None
        self._pack()

    def _storage_ids(self):
        return tuple(p.data_ptr() for p in self.params.values())

    # -- buffers -------------------------------------------------------------------------------------------
    def _buf(self, name: str, shape: Sequence[int], dtype) -> torch.Tensor:
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            for k in [k for k in self._ws if k[0] == name]:
                del self._ws[k]
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    def _timed(self, kind: str, flops: float, fn, *a, **k):
        """Launch through `fn`; when `self.timing` is a list, bracket the launch with CUDA events on the launching
        stream so bench.py can attribute device time per kernel family inside the timed region."""
        self.launches += 1
        if self.timing is None:
            return fn(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        self.timing.append((kind, flops, e0, e1))
        return r

    def _gemm(self, a, w, bias, out, epi, **k):
        return self._timed("gemm", 2.0 * a.shape[0] * w.shape[0] * a.shape[1], ops.gemm, a, w, bias, out, epi, **k)

    def _fmha(self, q, k, v, out, scale):
        B, Lq, H, hd = q.shape
        return self._timed("fmha", 4.0 * B * H * Lq * k.shape[1] * hd, ops.fmha, q, k, v, out, scale)

    # -- stages --------------------------------------------------------------------------------------------
    def _embed_mlp(self, pe: str, pp: str, values: torch.Tensor):
        """fp32 sinusoid -> Linear -> SiLU -> Linear (= e) -> SiLU -> Linear (= e0) on `values` [n]. (:928-955)"""
        P = self.params
        if values.numel() > 16:     # the weight-streaming kernel re-reads the weights every 4 rows: tensor cores instead
            return self._embed_mlp_tokens(values, pe, pp)
        emb = ops.sinusoid(values, self.cfg["freq_dim"])
        h = ops.linear_f32(emb, P[pe + ".0.weight"], P[pe + ".0.bias"], 0)
        e = ops.linear_f32(h, P[pe + ".2.weight"], P[pe + ".2.bias"], 1)
        e0 = ops.linear_f32(e, P[pp + ".1.weight"], P[pp + ".1.bias"], 1)
        self.launches += 4
        return e, e0

    def _cnn_fuser(self, y_ctrl0: torch.Tensor, add: torch.Tensor) -> torch.Tensor:
        """cnn_conv1..5 (:868-880) for one sample; inputs [48,F,H,W] and [240,F,H,W] bf16 -> channel-last [P, 48].

        The convolutions are per frame (kernel (1,3,3)), so across a sequence-parallel group every rank runs them on
        its own block of frames; only the GroupNorm statistics span the sample: per-(frame, group) fp64 partials are
        all-gathered and summed in frame order on every rank, the same order a single GPU uses, so the result is
        bit-identical on every layout. One more all-gather returns the 48-channel output to all ranks."""
        P = self.params
        C0, F, H, W = y_ctrl0.shape
        C1 = add.shape[0]
        pp = H * W
        par = self.par
        nsp = par.layout.sp_size if par is not None else 1
        Fc = -(-F // nsp)                                     # frames per rank; the last rank(s) may hold fewer or none
        f0 = min((par.layout.sp_rank if par is not None else 0) * Fc, F)
        Fl = min(f0 + Fc, F) - f0
        npix = Fl * pp
        out_pad = self._buf("cnn_out_pad", (Fc * pp, self.w_cnn5.shape[0]), bf16)
        if nsp > 1 and Fl < Fc:
            out_pad.zero_()
        # Activations of the 3x3 stages live in the zero-padded channel-last grid the implicit-GEMM convolution reads
        # ([Fl, H+2, W+2, C rounded up to 64]); only interior positions / real channels are ever written, so the halo and
        # the padding channels keep the zeros they were allocated with.
        act = None
        if Fl > 0:
            act = self._buf_zero("cnn_in_pad", (Fl * (H + 2) * (W + 2), -(-(C0 + C1) // 64) * 64), bf16)
            src0 = y_ctrl0 if nsp == 1 else y_ctrl0[:, f0:f0 + Fl].contiguous()
            src1 = add if nsp == 1 else add[:, f0:f0 + Fl].contiguous()
            ops.nchw_to_nhwc_padded(src0.reshape(C0, npix), act, 0, Fl, H, W)
            ops.nchw_to_nhwc_padded(src1.reshape(C1, npix), act, C0, Fl, H, W)
            self.launches += 2
        stats = self._buf("cnn_stats", (64,), f32)
        resid = None
        for j, (groups, keep) in enumerate([(24, True), (24, False), (12, True), (12, False)]):
            wj = self.w_cnn_taps[j]
            cout = wj.shape[0]
            part = self._buf(f"cnn_part{j}", (Fc, groups, 2), torch.float64)
            if Fl < Fc:
                part.zero_()
            if Fl > 0:
                conv = self._buf(f"cnn_conv{j}", (npix, cout), bf16)
                self._timed("gemm", 2.0 * npix * cout * wj.shape[1], ops.conv_gemm, act, wj,
                            P[f"cnn_conv{j + 1}.0.bias"], conv, Fl, H, W, 1, 3, FX_EPI_BF16)
                ops.groupnorm_partials(conv, Fl, groups, part[:Fl])
                self.launches += 1
            allp = part if nsp == 1 else par.gather_tokens(part)[:F].contiguous()     # [F, groups, 2], frame order
            if Fl > 0:
                last = j == 3        # x4 feeds the 1x1 convolution: dense rows
                nxt = (self._buf(f"cnn_act{j}", (npix, cout), bf16) if last else
                       self._buf_zero(f"cnn_act{j}_pad", (Fl * (H + 2) * (W + 2), -(-cout // 64) * 64), bf16))
                keep_f32 = self._buf(f"cnn_f32_{j}", (npix, cout), f32) if keep else None
                # x2 = S(GN(conv2(x1))) + x1 and x4 = S(GN(conv4(x3))) + x3 take the previous stage's fp32 output
                ops.groupnorm_silu_partials(conv, groups, 1e-5, P[f"cnn_conv{j + 1}.1.weight"],
                                            P[f"cnn_conv{j + 1}.1.bias"], allp, pp, None if keep else resid, keep_f32,
                                            nxt, stats, pad_hw=None if last else (H, W))
                self.launches += 2
                resid = keep_f32
                act = nxt
        if Fl > 0:
            self._gemm(act, self.w_cnn5, P["cnn_conv5.bias"], out_pad[:npix], FX_EPI_BF16)
        if nsp == 1:
            return out_pad.clone()
        return par.gather_tokens(out_pad)[:F * pp].contiguous()

    def _buf_zero(self, name: str, shape: Sequence[int], dtype) -> torch.Tensor:
        """A workspace that is zero when first handed out (halo / padding of the convolution grids)."""
        key = (name, tuple(shape), dtype)
        fresh = key not in self._ws
        t = self._buf(name, shape, dtype)
        if fresh:
            t.zero_()
        return t

    def _context(self, context: List[torch.Tensor]) -> torch.Tensor:
        """zero-pad to text_len, text_embedding MLP (:958-964) -> [B*text_len, D] bf16."""
        P, T = self.params, self.cfg["text_len"]
        B = len(context)
        padded = torch.zeros((B * T, self.cfg["text_dim"]), dtype=bf16, device=self.device)
        for b, u in enumerate(context):
            if u.shape[0] > T:
                raise FlexamNativeError(f"context {b} longer than text_len {T}")
            padded[b * T: b * T + u.shape[0]].copy_(u)
        h = torch.empty((B * T, self.D), dtype=bf16, device=self.device)
        self._gemm(padded, P["text_embedding.0.weight"], P["text_embedding.0.bias"], h, FX_EPI_GELU_BF16)
        ctx = torch.empty((B * T, self.D), dtype=bf16, device=self.device)
        self._gemm(h, P["text_embedding.2.weight"], P["text_embedding.2.bias"], ctx, FX_EPI_BF16)
        return ctx

    def _cross_kv(self, ctx: torch.Tensor) -> List[torch.Tensor]:
        """Per-layer cross-attention K (RMS-normed) | V of the context (:364-365): step-invariant."""
        out = []
        for w in self.blk:
            kv = torch.empty((ctx.shape[0], 2 * self.D), dtype=bf16, device=self.device)
            self._gemm(ctx, w["cwkv"], w["cbkv"], kv, FX_EPI_BF16)
            ops.rmsnorm_rope(kv[:, : self.D], w["cnk"], self.eps)
            self.launches += 1
            out.append(kv)
        return out

    @staticmethod
    def _ident(ts) -> tuple:
        return tuple((t.data_ptr(), t._version, tuple(t.shape), t.stride(), t.dtype) for t in ts)

    def _static_probe(self, srcs):
        """State of the step-invariant cache for these inputs: "hit" (the very same storages at the same version —
        sound only because the engine keeps references to the tensors it saw, so the caching allocator cannot recycle
        their addresses), "miss", or a 0-dim device tensor that is 1 when the CONTENTS equal the private copies (a
        fresh tensor with the same values, the sampler's per-step ``torch.cat``, still hits; read by _read_facts)."""
        key = self._static_key
        if key is None:
            return "miss"
        ident, refs, copies = key
        if ident == self._ident(srcs):
            return "hit"
        if self.trusted or len(copies) != len(srcs) or any(a.shape != b.shape for a, b in zip(copies, srcs)):
            return "miss"
        return torch.stack([(a == b).all() for a, b in zip(copies, srcs)]).all()

    def _read_facts(self, facts: dict) -> dict:
        """ONE device->host copy for every data-dependent fact of this call (0-dim device tensors -> Python ints)."""
        dev_items = [(k, v) for k, v in facts.items() if torch.is_tensor(v)]
        out = {k: v for k, v in facts.items() if not torch.is_tensor(v)}
        if dev_items:
            vals = torch.stack([v.reshape(()).to(torch.int64) for _, v in dev_items]).tolist()
            self.host_reads += 1
            out.update({k: int(x) for (k, _), x in zip(dev_items, vals)})
        return out

    # -- the denoising step ----------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, context, seq_len, y, full_ref, additional_control, density,
                block_hook=None, teacache=None, cond_flag=True, t_dedup=None, teacache_decision=None) -> torch.Tensor:
        """Returns the stacked prediction [B, out_dim, F, H, W] in bf16 (forward :817-1123).

        ``t_dedup=(uniq [U] f32, inv [B, L0] int32)``: the caller already knows the distinct per-token timesteps (the
        sampling loop: t = mask * t_step with a constant mask); ``teacache_decision``: the caller precomputed the
        TeaCache decision of this step (it depends on the timestep embedding only). With both, and ``trusted`` set,
        the call enqueues without a single host synchronisation (DenoiseLoop, CUDA-graph capturable)."""
        cfg, D, dev = self.cfg, self.D, self.device
        self.refresh_if_modified()
        self.launches = 0
        self.host_reads = 0
        if y is None or full_ref is None or additional_control is None or density is None:
            raise FlexamNativeError("the FlexAM path needs y, full_ref, additional_control and density")
        C = cfg["out_dim"]
        x = x.to(dev, bf16).contiguous()
        y = y.to(dev, bf16).contiguous()
        add = additional_control.to(dev, bf16).contiguous()
        full_ref = full_ref.to(dev, bf16).contiguous()
        context = [u.to(dev, bf16) for u in context]
        t = t.to(dev, f32)
        density = density.to(dev, f32)

        # ---- partitioning (flexam_b200/dist.py): CFG rows, then a token slice per sequence-parallel rank ------
        par = self.par
        B_all = x.shape[0]
        cfg_split = par is not None and par.layout.cfg_size > 1 and B_all % par.layout.cfg_size == 0
        row_lo = 0
        if cfg_split:
            nb = B_all // par.layout.cfg_size
            row_lo = par.layout.cfg_rank * nb
            x, y, add, full_ref, t, density = (u[row_lo:row_lo + nb] for u in (x, y, add, full_ref, t, density))
            context = context[row_lo:row_lo + nb]
            if t_dedup is not None:
                t_dedup = (t_dedup[0], t_dedup[1][row_lo:row_lo + nb])
        P = par.layout.sp_size if par is not None else 1
        sp_rank = par.layout.sp_rank if par is not None else 0

        B, _, F, Hh, Ww = x.shape
        Hp, Wp = Hh // 2, Ww // 2
        L0 = F * Hp * Wp
        R = Hp * Wp
        L = L0 + R
        if seq_len != L0:
            raise FlexamNativeError(f"seq_len {seq_len} != tokens on the grid {L0} (padding is handled by the SP layer)")
        if max(F + 1, Hp, Wp) > self.freqs.shape[0]:
            raise FlexamNativeError(f"latent grid ({F + 1},{Hp},{Wp}) exceeds the {self.freqs.shape[0]}-position RoPE table")
        grid = (F + 1, Hp, Wp)
        Lp = -(-L // P)            # tokens per SP rank; the tail of the last rank may be padding (:919-920)
        L_pad = Lp * P
        tok0 = sp_rank * Lp
        M = B * Lp

        # ---- data-dependent facts of this call, gathered with ONE device->host read ---------------------------
        # (a) weights edited behind the version counters (merge_lora's `weight.data += ...`): sampled fingerprint;
        # (b) step-invariant inputs unchanged (by content) -> reuse control fuser / context / cross K/V;
        # (c) number of distinct per-token timesteps -> modulation-table path or per-token path.
        facts = {}
        srcs = [y, add] + list(context)
        if not self.trusted and self._fp_table is not None:
            facts["weights_same"] = (ops.fingerprint(self._fp_table, self.FP_STRIDE) == self._fp_ref).all()
            self.launches += 1
        facts["static"] = self._static_probe(srcs) if self.cache_static else "miss"
        per_token = t.dim() == 2
        if per_token:
            if t.shape[1] < L:   # ref tokens are PREPENDED and take the last token's timestep (:900-904)
                t = torch.cat([t[:, -1:].expand(B, L - t.shape[1]), t], dim=1)
            t = t.contiguous()
            if t_dedup is not None:
                uniq, inv0 = t_dedup
                uniq = uniq.to(dev, f32).contiguous()
                inv0 = inv0.to(dev, i32)
                if inv0.shape[1] < L:
                    inv0 = torch.cat([inv0[:, -1:].expand(B, L - inv0.shape[1]), inv0], dim=1)
                inv = inv0.contiguous().view(-1)
                facts["n_uniq"] = int(uniq.numel())
            else:
                cap = self.max_table_timesteps
                uniq = self._buf("t_uniq", (cap,), f32)
                inv = self._buf("t_inv", (B * L,), i32)
                cnt = self._buf("t_count", (1,), i32)
                ops.dedup_f32(t.view(-1), cap, uniq, inv, cnt)
                self.launches += 1
                facts["n_uniq"] = cnt[0]
        facts = self._read_facts(facts)
        if not facts.get("weights_same", 1):
            self._pack()                   # re-derive the packed copies, drop everything cached from the old weights
            facts["static"] = "miss"
        static_hit = facts["static"] == "hit" or facts["static"] == 1

        # ---- step-invariant work: control fuser, context embedding, cross-attention K/V ----------------------
        if not static_hit:
            st = {}
            st["cnn"] = [self._cnn_fuser(y[b, :C], add[b]) for b in range(B)]
            st["ctx"] = self._context(context)
            st["kv"] = self._cross_kv(st["ctx"])
            self._static = st
            self._static_key = ((self._ident(srcs), list(srcs), [u.clone() for u in srcs])
                                if self.cache_static else None)
        elif facts["static"] == 1:         # same values at a new address: follow them
            self._static_key = (self._ident(srcs), list(srcs), self._static_key[2])
        st = self._static

        # ---- patch + ref embedding straight into the fp32 residual stream (:885-899) --------------------------
        xs = self._buf("x", (M, D), f32)
        x_full = xs if P == 1 else self._buf("x_full", (B * L_pad, D), f32)   # the embed is cheap: done for all tokens
        if L_pad != L:
            x_full.view(B, L_pad, D)[:, L:].zero_()
        rows = self._buf("patch_rows", (L0, self.w_patch.shape[1]), bf16)
        rrows = self._buf("ref_rows", (R, self.w_ref.shape[1]), bf16)
        for b in range(B):
            o = b * L_pad
            ops.patchify([x[b], st["cnn"][b].view(F, Hh, Ww, -1), y[b, C:]], [False, True, False], F, Hh, Ww, rows)
            self._gemm(rows, self.w_patch, self.params["patch_embedding.bias"], x_full[o + R: o + L], FX_EPI_F32)
            ops.patchify([full_ref[b].unsqueeze(1)], [False], 1, Hh, Ww, rrows)
            self._gemm(rrows, self.w_ref, self.params["ref_conv.bias"], x_full[o: o + R], FX_EPI_F32)
            self.launches += 2
        if P > 1:   # keep this rank's token slice (torch.chunk(x, P, dim=1)[rank], :971-975)
            xs.view(B, Lp, D).copy_(x_full.view(B, L_pad, D)[:, tok0:tok0 + Lp])

        # ---- timestep / density embeddings (:900-955) --------------------------------------------------------
        # table mode: the fp32 MLPs run on the DISTINCT timesteps and the block LayerNorms read rows combined once per
        #             (timestep, sample); token mode (many distinct per-token values): the MLPs run per token of this
        #             rank's slice on the tensor cores and LayerNorm / gates read the per-token e0 (:444-446).
        token_mode = per_token and facts["n_uniq"] > self.max_table_timesteps
        de, de0 = self._embed_mlp("density_embedding", "density_projection", density.contiguous())
        de0v = de0.view(B, 2, D)
        if token_mode:
            t_pad = t if L_pad == L else torch.cat([t, t[:, -1:].expand(B, L_pad - L)], dim=1)   # :931-935
            t_loc = t_pad[:, tok0:tok0 + Lp].contiguous().view(-1)
            e, e0 = self._embed_mlp_tokens(t_loc)                                   # [M, D], [M, 6D]
            row_idx = self._arange(M)
            e_last, e0_last = self._embed_mlp("time_embedding", "time_projection", t[:, -1].contiguous())
            mod_inp = e0_last.view(B, 6, D)
        else:
            if per_token:
                U = facts["n_uniq"]
                uniq = uniq[:U]
                idx_full = inv.view(B, L)
            else:
                uniq = t.contiguous()
                U = uniq.shape[0]
                idx_full = torch.arange(B, device=dev, dtype=i32).view(B, 1).expand(B, L)
            last_idx = idx_full[:, -1].long()
            if L_pad != L:           # padding rows reuse the last token's timestep (:931-935)
                idx_full = torch.cat([idx_full, idx_full[:, -1:].expand(B, L_pad - L)], dim=1)
            row_idx = idx_full[:, tok0:tok0 + Lp].contiguous().view(-1)
            e, e0 = self._embed_mlp("time_embedding", "time_projection", uniq.contiguous())      # [U,D], [U,6D]
            mod_inp = None
        e0v = e0.view(-1, 6, D)

        # ---- 30 x WanAttentionBlock (:422-472) ---------------------------------------------------------------
        h = self._buf("h", (M, D), bf16)
        qkv = self._buf("qkv", (M, 3 * D), bf16)
        fused_sp = P > 1 and par.fused
        if fused_sp:    # the o-projection input is written by the peers: symmetric memory
            attn_sym = par.symm_buffer("attn", (B, Lp, D), bf16, dev)
            attn = attn_sym.tensor.view(M, D)
        else:
            attn = self._buf("attn", (M, D), bf16)
        cq = self._buf("cq", (M, D), bf16)
        ffn = self._buf("ffn", (M, cfg["ffn_dim"]), bf16)
        T = cfg["text_len"]
        scale = 1.0 / math.sqrt(128.0)
        qkv5 = qkv.view(B, Lp, 3, self.H, 128)
        attn4 = attn.view(B, Lp, self.H, 128)
        cq4 = cq.view(B, Lp, self.H, 128)
        if not token_mode:
            # the block LayerNorms read modulation rows combined once per (distinct timestep, sample): row u*B + b
            row_idx2 = (row_idx.view(B, Lp) * B + torch.arange(B, device=dev, dtype=i32).view(B, 1)).view(-1)
            modtab = self._buf("modtab", (2, U * B, 2, D), f32)

        # ---- TeaCache (:977-1051): skip the block stack and re-apply the previous residual when the modulated
        #      timestep embedding moved little. The decision uses the GLOBAL last token so all SP ranks agree.
        run_blocks = True
        ori = None
        if teacache is not None:
            if teacache_decision is not None:            # precomputed by the sampling loop (depends on t only)
                run_blocks = bool(teacache_decision)
                if cond_flag:
                    teacache.should_calc = run_blocks
            else:
                # this package's TeaCache has decide()/step_done(); the reference's own class (an installed module whose
                # pipeline called the reference's enable_teacache) has the same state and is driven through the functions
                inp = mod_inp if mod_inp is not None else e0v[last_idx]
                run_blocks = (teacache.decide(inp, cond_flag) if hasattr(teacache, "decide")
                              else _tc.decide(teacache, inp, cond_flag))
                self.host_reads += 1 if cond_flag else 0
            if not run_blocks:
                prev = self._teacache_residual(teacache, cond_flag, B, M, cfg_split)
                ops.add_(xs, prev)
                self.launches += 1
            else:
                ori = xs.clone()
        for i, w in enumerate(self.blk if run_blocks else ()):
            mod, dmod = w["mod"], w["dmod"]
            # self-attention
            if token_mode:
                ops.ln_modulate(xs, h, self.eps, mod[0], mod[1], e0v[:, 0], e0v[:, 1], 6 * D, row_idx, dmod[0],
                                de0v[:, 0], 2 * D, Lp)
            else:
                ops.modulation_tables(mod, dmod, e0v, de0v, modtab)
                ops.ln_scale_shift(xs, h, self.eps, modtab[0, :, 0], modtab[0, :, 1], 2 * D, row_idx2)
            self._gemm(h, w["wqkv"], w["bqkv"], qkv, FX_EPI_BF16)
            if fused_sp:    # Ulysses with the exchange fused into the norm/rope and attention kernels (peer stores)
                par.attention_fused(qkv, attn_sym, w["nq"], w["nk"], self.eps, self.freqs, grid, L, scale, ops,
                                    timed=self._timed)
            else:
                ops.rmsnorm_rope(qkv[:, :2 * D], w["nq"], self.eps, self.freqs, grid, tok0, Lp, weight2=w["nk"])
                if P == 1:
                    self._fmha(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], attn4, scale)
                else:       # Ulysses through NCCL all-to-alls, one sample at a time
                    for b in range(B):
                        par.attention(qkv5[b], attn4[b], L, scale)
            self._gemm(attn, w["wo"], w["bo"], xs, FX_EPI_RESID_F32, gate_mod=mod[2], gate_e=e0v[:, 2], row_idx=row_idx)
            # cross-attention (no gate, no RoPE, all text_len slots attended)
            ops.ln_affine(xs, h, self.eps, w["n3w"], w["n3b"])
            self._gemm(h, w["cwq"], w["cbq"], cq, FX_EPI_BF16)
            ops.rmsnorm_rope(cq, w["cnq"], self.eps)
            kv5 = st["kv"][i].view(B, T, 2, self.H, 128)
            self._fmha(cq4, kv5[:, :, 0], kv5[:, :, 1], attn4, scale)
            self._gemm(attn, w["cwo"], w["cbo"], xs, FX_EPI_RESID_F32)
            # ffn
            if token_mode:
                ops.ln_modulate(xs, h, self.eps, mod[3], mod[4], e0v[:, 3], e0v[:, 4], 6 * D, row_idx, dmod[1],
                                de0v[:, 1], 2 * D, Lp)
            else:
                ops.ln_scale_shift(xs, h, self.eps, modtab[1, :, 0], modtab[1, :, 1], 2 * D, row_idx2)
            self._gemm(h, w["w1"], w["b1"], ffn, FX_EPI_GELU_BF16)
            self._gemm(ffn, w["w2"], w["b2"], xs, FX_EPI_RESID_F32, gate_mod=mod[5], gate_e=e0v[:, 5], row_idx=row_idx)
            self.launches += 5 if token_mode else 6
            if block_hook is not None:
                block_hook(i, xs)
        if teacache is not None and run_blocks:
            # a persistent buffer per branch: a CUDA-graph replay of this call and of a later skipped step then agree on
            # where the residual lives (the reference's `offload` to host memory is not needed with 180 GB of HBM)
            res = self._buf("tc_res_cond" if cond_flag else "tc_res_uncond", tuple(xs.shape), f32)
            ops.sub(xs, ori, res)
            self.launches += 1
            if cond_flag:
                teacache.previous_residual_cond = res
            else:
                teacache.previous_residual_uncond = res
            # which rows of the caller's batch this residual covers (cfg-parallel ranks hold one branch each)
            teacache._fx_rows = (row_lo, row_lo + B, B_all if cfg_split else B)

        # ---- head (:493-507, uses e not e0) + unpatchify (:1106-1149) -----------------------------------------
        ops.ln_modulate(xs, h, self.eps, self.head_mod[0], self.head_mod[1], e, e, D, row_idx, self.head_dmod, de, D, Lp)
        ho = self._buf("head", (M, self.params["head.head.weight"].shape[0]), bf16)
        self._gemm(h, self.params["head.head.weight"], self.params["head.head.bias"], ho, FX_EPI_BF16)
        out = torch.empty((B, C, F, Hh, Ww), dtype=bf16, device=dev)
        for b in range(B):
            tokens = ho[b * Lp: (b + 1) * Lp] if P == 1 else par.gather_tokens(ho[b * Lp: (b + 1) * Lp])  # :1103-1104
            ops.unpatchify(tokens[R:L], out[b])     # the ref tokens are dropped from the FRONT (:1106-1109)
        self.launches += 1 + B
        if cfg_split:
            out = par.gather_cfg(out)
        if teacache is not None:
            teacache.step_done(cond_flag) if hasattr(teacache, "step_done") else _tc.step_done(teacache, cond_flag)
        return out

    def _teacache_residual(self, teacache, cond_flag: bool, B: int, M: int, cfg_split: bool) -> torch.Tensor:
        """The residual the reference adds on a skipped step: ``previous_residual[-x.size(0):]`` (:1003-1006), i.e. the
        LAST B rows of the batch that was stored. Under CFG-branch parallelism every rank stored only its own branch
        (cfg_rank 0 = uncond); when the batch later shrinks to the cond half (cfg_skip) the split is off and every rank
        needs the COND rows, which live on the last cfg rank: fetched once from that peer (all ranks take this path
        together, the decision depends on the timestep only), then kept locally."""
        prev = teacache.previous_residual_cond if cond_flag else teacache.previous_residual_uncond
        lo, hi, total = getattr(teacache, "_fx_rows", (0, prev.shape[0] // max(M // B, 1), prev.shape[0] // max(M // B, 1)))
        stored_split = total > hi - lo
        if stored_split and not cfg_split:
            par = self.par
            if par is None or B > hi - lo:
                raise FlexamNativeError("TeaCache: cached residual was stored under a CFG-parallel layout that does not "
                                        "cover the rows needed now; disable TeaCache or keep the layout fixed")
            prev = par.fetch_from_last_cfg_rank(prev.contiguous())
            if cond_flag:
                teacache.previous_residual_cond = prev
            else:
                teacache.previous_residual_uncond = prev
            teacache._fx_rows = (total - (hi - lo), total, hi - lo)     # now a local, unsplit record of the last rows
        return prev[-M:].contiguous()

    def _arange(self, n: int) -> torch.Tensor:
        t = self._ws.get(("arange", n))
        if t is None:
            t = torch.arange(n, device=self.device, dtype=i32)
            self._ws[("arange", n)] = t
        return t

    def _embed_mlp_tokens(self, t_rows: torch.Tensor, pe: str = "time_embedding", pp: str = "time_projection"):
        """time_embedding / time_projection (:928-944) for MANY rows (per-token timesteps that do not de-duplicate):
        fp32 inputs split into bf16 planes, contractions on tcgen05 (fx_linear_f32_tc). t_rows: [M] f32 ->
        e [M, D], e0 [M, 6D] fp32."""
        P, D, Mr = self.params, self.D, t_rows.numel()
        emb = ops.sinusoid(t_rows, self.cfg["freq_dim"])
        ws = self._buf("mlp_planes", (Mr * self.mlp_planes * max(D, self.cfg["freq_dim"]),), bf16)
        h1 = ops.linear_f32_tc(emb, P[pe + ".0.weight"], P[pe + ".0.bias"], 0, self.mlp_planes,
                               self._buf("mlp_h1", (Mr, D), f32), ws)
        e = ops.linear_f32_tc(h1, P[pe + ".2.weight"], P[pe + ".2.bias"], 1, self.mlp_planes,
                              self._buf("mlp_e", (Mr, D), f32), ws)
        w3 = P[pp + ".1.weight"]
        e0 = ops.linear_f32_tc(e, w3, P[pp + ".1.bias"], 1, self.mlp_planes,
                               self._buf("mlp_e0", (Mr, w3.shape[0]), f32), ws)
        self.launches += 7
        return e, e0

    def _swap01(self, src, out):
        self.launches += 1
        return ops.swap01(src, out)


# ----------------------------------------------------------------------------------------------------------
# nn.Module with the reference's surface
# ----------------------------------------------------------------------------------------------------------
def _set_param(root: nn.Module, dotted: str, p: nn.Parameter):
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        child = m._modules.get(name)
        if child is None:
            child = nn.Module()
            m.add_module(name, child)   # digit names ("0", "2") mirror the reference's nn.Sequential children
        m = child
    m.register_parameter(parts[-1], p)


class _Config(dict):
    __getattr__ = dict.get


class Wan2_2Transformer3DModel_FlexAM(nn.Module):
    """Drop-in for FlexAM.models.Wan2_2Transformer3DModel_FlexAM (:1335-1438) with a native forward."""

    def __init__(self, model_type="t2v", patch_size=(1, 2, 2), text_len=512, in_dim=16, dim=2048, ffn_dim=8192,
                 freq_dim=256, text_dim=4096, out_dim=16, num_heads=16, num_layers=32, window_size=(-1, -1),
                 qk_norm=True, cross_attn_norm=True, eps=1e-6, in_channels=16, hidden_size=2048,
                 add_control_adapter=False, in_dim_control_adapter=24, downscale_factor_control_adapter=8,
                 add_ref_conv=False, in_dim_ref_conv=16, add_cnn_block=False, in_dim_cnn_block=96,
                 out_dim_cnn_block=16, dtype=torch.bfloat16, device=None):
        super().__init__()
        if not (add_ref_conv and add_cnn_block and qk_norm and cross_attn_norm) or add_control_adapter:
            raise FlexamNativeError("native FlexAM path: add_ref_conv, add_cnn_block, qk_norm, cross_attn_norm must be "
                                    "on and add_control_adapter off (config/wan2.2/wan_civitai_5b_FlexAM.yaml)")
        if tuple(window_size) != (-1, -1) or tuple(patch_size) != (1, 2, 2):
            raise FlexamNativeError("native FlexAM path: global attention and patch (1,2,2) only")
        self.config = _Config(
            model_type=model_type, patch_size=tuple(patch_size), text_len=text_len, in_dim=in_dim, dim=dim,
            ffn_dim=ffn_dim, freq_dim=freq_dim, text_dim=text_dim, out_dim=out_dim, num_heads=num_heads,
            num_layers=num_layers, window_size=tuple(window_size), qk_norm=qk_norm, cross_attn_norm=cross_attn_norm,
            eps=eps, in_channels=in_channels, hidden_size=hidden_size, add_control_adapter=add_control_adapter,
            add_ref_conv=add_ref_conv, in_dim_ref_conv=in_dim_ref_conv, add_cnn_block=add_cnn_block,
            in_dim_cnn_block=in_dim_cnn_block, out_dim_cnn_block=out_dim_cnn_block)
        for k in ("model_type", "patch_size", "text_len", "in_dim", "dim", "ffn_dim", "freq_dim", "text_dim", "out_dim",
                  "num_heads", "num_layers", "eps"):
            setattr(self, k, self.config[k])
        self.blocks = nn.ModuleList([nn.Module() for _ in range(num_layers)])
        for name, shape in param_shapes(self.config).items():
            _set_param(self, name, nn.Parameter(torch.empty(shape, dtype=dtype, device=device), requires_grad=False))
        self.d = dim // num_heads
        self.teacache = None
        self.cfg_skip_ratio = None
        self.current_steps = 0
        self.num_inference_steps = None
        self.gradient_checkpointing = False
        self.sp_world_size = 1
        self.sp_world_rank = 0
        self.freqs = rope_complex(self.d)          # plain attribute like the reference's (callers read and re-assign it)
        self._engine: Optional[NativeEngine] = None

    # diffusers' ModelMixin conveniences the pipelines rely on
    @property
    def dtype(self) -> torch.dtype:
        return next(self.parameters()).dtype

    @property
    def device(self) -> torch.device:
        return next(self.parameters()).device

    # -- loading (:1190-1332) ------------------------------------------------------------------------------------
    @classmethod
    def from_config(cls, config: dict, **kwargs):
        """Constructor arguments from a diffusers-style config dict (unknown keys such as ``_class_name`` ignored), like
        ``ConfigMixin.from_config`` as the reference uses it (:1227, :1289)."""
        import inspect
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        merged = {k: v for k, v in dict(config, **kwargs).items() if k in accepted}
        return cls(**merged)

    @classmethod
    def from_pretrained(cls, pretrained_model_path, subfolder=None, transformer_additional_kwargs={},
                        low_cpu_mem_usage=False, torch_dtype=torch.bfloat16):
        """Same contract as the reference (:1190-1332): ``config.json`` + ``diffusion_pytorch_model.{bin,safetensors}``
        or ``*.safetensors`` shards under ``path[/subfolder]``; ``dict_mapping``; a checkpoint ``patch_embedding.weight``
        with fewer / more input channels is copied into the leading channels and the rest zero-filled; tensors whose size
        does not match are skipped; non-strict load; cast to ``torch_dtype``. ``low_cpu_mem_usage`` is accepted and gives
        the same result (the reference's meta-device path is an optimisation of the same load)."""
        import glob
        import json
        import os
        if subfolder is not None:
            pretrained_model_path = os.path.join(pretrained_model_path, subfolder)
        config_file = os.path.join(pretrained_model_path, "config.json")
        if not os.path.isfile(config_file):
            raise RuntimeError(f"{config_file} does not exist")
        with open(config_file, "r") as f:
            config = json.load(f)
        kwargs = dict(transformer_additional_kwargs)
        for key, target in dict(kwargs.pop("dict_mapping", {})).items():
            kwargs[target] = config[key]
        model = cls.from_config(config, **kwargs)
        model_file = os.path.join(pretrained_model_path, "diffusion_pytorch_model.bin")      # diffusers WEIGHTS_NAME
        model_file_safetensors = model_file.replace(".bin", ".safetensors")
        if os.path.exists(model_file):
            state_dict = torch.load(model_file, map_location="cpu")
        else:
            from safetensors.torch import load_file
            files = ([model_file_safetensors] if os.path.exists(model_file_safetensors)
                     else sorted(glob.glob(os.path.join(pretrained_model_path, "*.safetensors"))))
            state_dict = {}
            for fn in files:
                state_dict.update(load_file(fn))
        own = model.state_dict()
        pe = "patch_embedding.weight"
        if pe in state_dict and own[pe].size() != state_dict[pe].size():
            n_ckpt, n_own = state_dict[pe].size(1), own[pe].size(1)
            w = torch.zeros_like(own[pe])
            w[:, :min(n_ckpt, n_own)] = state_dict[pe][:, :min(n_ckpt, n_own)].to(w.dtype)
            state_dict[pe] = w
        kept = {}
        for key, val in state_dict.items():
            if key in own and own[key].size() == val.size():
                kept[key] = val
            else:
                print(key, "Size don't match, skip")
        missing, unexpected = model.load_state_dict(kept, strict=False)
        print(f"### missing keys: {len(missing)}; \n### unexpected keys: {len(unexpected)};")
        if missing:     # the reference constructor ran init_weights() (:1151-1188); torch.empty storage is not a value
            model.init_missing_(missing)
            print("### initialised like the reference's init_weights (not in the checkpoint): " + ", ".join(missing))
        return model.to(torch_dtype)

    @torch.no_grad()
    def init_missing_(self, keys) -> None:
        """Values the reference gives parameters that no checkpoint tensor overwrote (init_weights :1151-1188 and the
        module constructors): Linear weights Xavier-uniform with zero bias, text / time embedding weights N(0, 0.02),
        density MLPs and ``head.head.weight`` ZERO (a base Wan2.2 checkpoint therefore loads as a no-op for the FlexAM
        additions), modulation rows N(0,1)/sqrt(dim) (:419-420, :490-491), norm weights 1 / biases 0, convolutions with
        torch's default (Kaiming-uniform) initialisation."""
        own = dict(self.named_parameters())
        D = self.config["dim"]
        for k in keys:
            p = own[k]
            leaf = k.rsplit(".", 1)[-1]
            if k.startswith(("density_embedding", "density_projection")) or k == "head.head.weight":
                p.zero_()
            elif "modulation" in leaf:
                p.copy_(torch.randn(p.shape) / D ** 0.5)
            elif "norm" in k and leaf == "weight" or (k.startswith("cnn_conv") and ".1.weight" in k):
                p.fill_(1.0)
            elif leaf == "bias" and ("conv" not in k or ".1.bias" in k):
                p.zero_()
            elif k.startswith(("text_embedding", "time_embedding")) and leaf == "weight":
                p.copy_(torch.randn(p.shape) * 0.02)
            elif p.dim() == 2:                                   # nn.Linear
                w = torch.empty(p.shape)
                nn.init.xavier_uniform_(w)
                p.copy_(w)
            elif leaf == "weight":                               # patch_embedding / ref_conv / cnn convolutions
                w = torch.empty(p.shape)
                if k == "patch_embedding.weight":
                    nn.init.xavier_uniform_(w.flatten(1))
                else:
                    nn.init.kaiming_uniform_(w, a=math.sqrt(5))
                p.copy_(w)
            else:                                                # conv bias: U(-1/sqrt(fan_in), 1/sqrt(fan_in))
                fan_in = own[k[:-4] + "weight"][0].numel()
                p.copy_((torch.rand(p.shape) * 2 - 1) / math.sqrt(fan_in))

    # -- feature toggles with the reference's names (:730-815) -----------------------------------------------
    def enable_cfg_skip(self, cfg_skip_ratio, num_steps):
        self.cfg_skip_ratio = cfg_skip_ratio if cfg_skip_ratio != 0 else None
        self.current_steps = 0
        self.num_inference_steps = num_steps if cfg_skip_ratio != 0 else None

    def share_cfg_skip(self, transformer=None):
        self.cfg_skip_ratio = transformer.cfg_skip_ratio
        self.current_steps = transformer.current_steps
        self.num_inference_steps = transformer.num_inference_steps

    def disable_cfg_skip(self):
        self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = None, 0, None

    def enable_riflex(self, k=6, L_test=66, L_test_scale=4.886):
        """Replaces ``self.freqs`` like the reference (:774-788); the engine's table follows it at the next forward."""
        self.freqs = rope_complex(self.d, riflex=dict(k=k, L_test=L_test, L_test_scale=L_test_scale)).to(self.freqs.device)

    def disable_riflex(self):
        self.freqs = rope_complex(self.d).to(self.freqs.device)

    def enable_teacache(self, coefficients, num_steps, rel_l1_thresh, num_skip_start_steps=0, offload=True):
        from .teacache import TeaCache
        self.teacache = TeaCache(coefficients, num_steps, rel_l1_thresh, num_skip_start_steps, offload)

    def share_teacache(self, transformer=None):
        self.teacache = transformer.teacache

    def disable_teacache(self):
        self.teacache = None

    def enable_multi_gpus_inference(self):
        from . import dist
        dist.attach(self)

    # -- forward -----------------------------------------------------------------------------------------------
    def engine(self) -> NativeEngine:
        if self._engine is None:
            params = {k: v.detach() for k, v in self.named_parameters()}   # detach() shares the version counter
            dev = next(iter(params.values())).device
            self._engine = NativeEngine(params, self.config, dev)
        return self._engine

    def forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None, subject_ref=None,
                cond_flag=True, additional_control=None, density=None):
        return native_forward(self, x, t, context, seq_len, clip_fea, y, y_camera, full_ref, subject_ref, cond_flag,
                              additional_control, density)


def native_forward(self, x, t, context, seq_len, clip_fea=None, y=None, y_camera=None, full_ref=None,
                   subject_ref=None, cond_flag=True, additional_control=None, density=None):
    """forward() of the reference (:817-1123) incl. the @cfg_skip wrapper (FlexAM/utils/cfg_optimization.py:5-38)."""
    if clip_fea is not None or y_camera is not None or subject_ref is not None:
        raise FlexamNativeError("clip_fea / y_camera / subject_ref are not part of the FlexAM path")
    bs = len(x)
    skip = (bs >= 2 and self.cfg_skip_ratio is not None and
            self.current_steps >= self.num_inference_steps * (1 - self.cfg_skip_ratio))
    if skip:  # run the cond half only and duplicate it, as the wrapper does
        h = bs // 2
        x, t, context, y, full_ref, additional_control, density = (
            x[h:], t[h:], context[h:], y[h:], full_ref[h:], additional_control[h:], density[h:])
    loop_kw = dict(getattr(self, "_fx_loop_kwargs", None) or {})    # DenoiseLoop: known timestep structure / decisions
    if skip and loop_kw.get("t_dedup") is not None:
        loop_kw["t_dedup"] = (loop_kw["t_dedup"][0], loop_kw["t_dedup"][1][bs // 2:])
    eng = self.engine() if hasattr(self, "engine") else self._flexam_engine
    # parameters whose storage was replaced since the engine saw them (module.to(...), `param.data = ...`): adopt them
    if eng._storage_ids() != tuple(p.data_ptr() for p in self.parameters()):
        eng.refresh({k: v.detach() for k, v in self.named_parameters()})
    _sync_rope_table(self, eng)          # the module's `freqs` attribute is the source of truth (enable_riflex() etc.)
    # the kernels launch on the calling thread's CURRENT device and on torch's current stream of that device: make the
    # module's device current for the whole call (a module on cuda:1 called while cuda:0 is current)
    guard = torch.cuda.device(eng.device) if eng.device.type == "cuda" else contextlib.nullcontext()
    with guard, ops.stream_scope():
        out = eng.forward(x, t, context, seq_len, y, full_ref, additional_control, density,
                          teacache=getattr(self, "teacache", None), cond_flag=cond_flag,
                          **loop_kw)
    if skip:
        out = torch.cat([out, out], dim=0)
    return out


def _sync_rope_table(module: nn.Module, eng: NativeEngine) -> None:
    """Keep the engine's (cos, sin) table equal to the reference module's complex ``freqs`` buffer (:655-665), which
    its own ``enable_riflex`` / ``disable_riflex`` (:774-799) replace at any time. Converted once per distinct buffer."""
    fr = getattr(module, "freqs", None)
    if not torch.is_tensor(fr) or not fr.is_complex():
        return
    key = (fr.data_ptr(), fr._version, tuple(fr.shape))
    if getattr(eng, "_freqs_key", None) == key:
        return
    if fr.dim() != 2 or tuple(fr.shape) != tuple(eng.freqs.shape[:2]):
        raise FlexamNativeError(f"module.freqs has shape {tuple(fr.shape)}, expected {tuple(eng.freqs.shape[:2])}")
    eng.freqs = torch.view_as_real(fr.detach().to("cpu", torch.complex128)).to(f32).contiguous().to(eng.device)
    eng._freqs_key = key
    eng._freqs_ref = fr          # keeps the storage alive so the key cannot be recycled


def refresh(module: nn.Module) -> None:
    """Tell the native engine of ``module`` (a mirror instance or an ``install``-ed reference module) that weights were
    edited behind autograd's back (``param.data += ...``, e.g. LoRA merge / unmerge): re-packs the copied weights."""
    eng = module._engine if hasattr(module, "engine") else getattr(module, "_flexam_engine", None)
    if eng is not None:
        eng.refresh({k: v.detach() for k, v in module.named_parameters()})


def install(module: nn.Module) -> nn.Module:
    """Rebind ``module.forward`` (a reference ``Wan2_2Transformer3DModel_FlexAM`` instance, bf16, on a B200) to the
    native path — the same method-rebinding plug-in pattern the reference uses for USP (:807-815).
    ``FLEXAM_BACKEND=reference`` turns the call into a no-op (the module keeps its torch forward); anything other than
    ``native`` / ``reference`` raises. There is no silent fallback: with ``native`` a missing library or a CPU module
    raises."""
    import os
    import types
    backend = os.environ.get("FLEXAM_BACKEND", "native")
    if backend == "reference":   # SURVEY §8b env switch: leave the module's own torch forward in place
        return module
    if backend != "native":
        raise FlexamNativeError(f"FLEXAM_BACKEND must be 'native' or 'reference', got {backend!r}")
    params = {k: v.detach() for k, v in module.named_parameters()}
    cfg = dict(module.config)
    cfg.setdefault("in_dim_ref_conv", params["ref_conv.weight"].shape[1])
    cfg.setdefault("in_dim_cnn_block", params["cnn_conv1.0.weight"].shape[1])
    cfg.setdefault("out_dim_cnn_block", params["cnn_conv5.weight"].shape[0])
    module._flexam_engine = NativeEngine(params, cfg, next(iter(params.values())).device)
    module.forward = types.MethodType(native_forward, module)
    return module
