"""fp32 verification mode of the native step: same launch sequence as ``NativeEngine.forward`` with fp32 activations.

Purpose: the north-star check "<= 1e-4 relative L2 against an fp32 run" of the reference
(FlexAM/models/wan_transformer3d_FlexAM.py:817-1123 executed in fp32). It is NOT the product path — the bf16
engine in ``model.py`` is — but it runs on the same C-ABI library and keeps every contraction on tcgen05:

* an fp32 activation matrix is split exactly into three bf16 planes (``fx_split3_f32``: a = hi + mid + lo), each
  plane goes through ``fx_gemm_bf16`` with the exact fp32 epilogue (``FX_EPI_F32_EXACT``) from the smallest plane up,
  and the fp32 results are added (``fx_add_f32``). Weights are bf16 parameters, so they need no split.
* gathers (patchify / im2col / unpatchify) move fp32 data losslessly plane by plane through the bf16 kernels.
* LayerNorm+modulation, RMSNorm+RoPE, GELU, gated residuals, GroupNorm and attention run as fp32 SIMT kernels
  (``csrc/precise.cu``) without the bf16 roundings of the autocast flow.

Single GPU, no TeaCache, no sequence parallelism: it exists to be compared with the fp32 goldens.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch

from . import ops
from .lib import FX_EPI_F32_EXACT, FlexamNativeError
from .model import NativeEngine

bf16, f32, i32 = torch.bfloat16, torch.float32, torch.int32


class PreciseEngine(NativeEngine):
    """NativeEngine with fp32 activations. ``k_chunk``: split long reductions into separately accumulated chunks
    whose results are added in fp32 round-to-nearest (None = one tcgen05 accumulation over the whole K). The tcgen05
    fp32 accumulator does not round to nearest: one accumulation over K = 14336 measured 1.5e-5 relative L2 against
    fp64 on B200 (tests/test_native_gpu.py), so the verification mode chunks by default."""

    k_chunk: Optional[int] = 1024

    # -- linear on exact bf16 planes ---------------------------------------------------------------------------
    def _linear_planes(self, planes: List[torch.Tensor], w: torch.Tensor, bias, out: torch.Tensor) -> torch.Tensor:
        """out f32 [M, N] = (sum_i planes[i]) @ w^T + bias; planes ordered most-significant first."""
        M, K = planes[0].shape
        kc = self.k_chunk or K
        jobs = []
        for pl in reversed(planes):                       # smallest contribution first
            for c0 in range(0, K, kc):
                jobs.append((pl[:, c0:c0 + kc], w[:, c0:c0 + kc]))
        tmp = None
        for j, (a, wk) in enumerate(jobs):
            last = j == len(jobs) - 1
            if j == 0:
                self._gemm(a, wk, bias if last else None, out, FX_EPI_F32_EXACT)
            else:
                if tmp is None:
                    tmp = self._buf("lin_tmp", tuple(out.shape), f32)
                self._gemm(a, wk, bias if last else None, tmp, FX_EPI_F32_EXACT)
                ops.add_(out, tmp)
                self.launches += 1
        return out

    def _split(self, a: torch.Tensor, name: str) -> List[torch.Tensor]:
        M, K = a.shape
        planes = self._buf("split_" + name, (3, M, K), bf16)
        ops.split3(a, planes)
        self.launches += 1
        return [planes[0], planes[1], planes[2]]

    def _linear(self, a: torch.Tensor, w: torch.Tensor, bias, out: torch.Tensor, name: str = "a") -> torch.Tensor:
        return self._linear_planes(self._split(a, name), w, bias, out)

    # -- stages ------------------------------------------------------------------------------------------------
    def _cnn_fuser_f32(self, y_ctrl0: torch.Tensor, add: torch.Tensor) -> torch.Tensor:
        """cnn_conv1..5 (:868-880) for one sample in fp32; returns channel-last f32 [P, 48]."""
        P = self.params
        C0, F, H, W = y_ctrl0.shape
        C1 = add.shape[0]
        npix = F * H * W
        act0 = self._buf("cnn_in", (npix, C0 + C1), bf16)       # inputs are bf16 tensors: one exact plane
        ops.nchw_to_nhwc(y_ctrl0.reshape(C0, npix), act0, 0)
        ops.nchw_to_nhwc(add.reshape(C1, npix), act0, C0)
        self.launches += 2
        planes = [act0]
        stats = self._buf("cnn_stats", (64,), f32)
        resid = None
        act = None
        for j, (groups, keep) in enumerate([(24, True), (24, False), (12, True), (12, False)]):
            wj = self.w_cnn[j]
            cin, cout = planes[0].shape[1], wj.shape[0]
            rows = []
            for i, pl in enumerate(planes):
                r = self._buf(f"cnn_rows{i}", (npix, 9 * cin), bf16)
                ops.im2col3x3(pl, F, H, W, r)
                self.launches += 1
                rows.append(r)
            conv = self._buf(f"cnn_conv_f32_{j}", (npix, cout), f32)
            self._linear_planes(rows, wj, P[f"cnn_conv{j + 1}.0.bias"], conv)
            act = self._buf(f"cnn_act_f32_{j}", (npix, cout), f32)
            ops.groupnorm_silu_f32(conv, groups, 1e-5, P[f"cnn_conv{j + 1}.1.weight"], P[f"cnn_conv{j + 1}.1.bias"],
                                   None if keep else resid, act, stats)
            self.launches += 2
            resid = act if keep else None
            planes = self._split(act, f"cnn{j}")
        out = torch.empty((npix, self.w_cnn5.shape[0]), dtype=f32, device=self.device)
        self._linear_planes(planes, self.w_cnn5, P["cnn_conv5.bias"], out)
        return out

    def _context_f32(self, context: List[torch.Tensor]) -> torch.Tensor:
        P, T = self.params, self.cfg["text_len"]
        B = len(context)
        padded = torch.zeros((B * T, self.cfg["text_dim"]), dtype=bf16, device=self.device)
        for b, u in enumerate(context):
            if u.shape[0] > T:
                raise FlexamNativeError(f"context {b} longer than text_len {T}")
            padded[b * T: b * T + u.shape[0]].copy_(u)
        h = torch.empty((B * T, self.D), dtype=f32, device=self.device)
        self._linear_planes([padded], P["text_embedding.0.weight"], P["text_embedding.0.bias"], h)
        ops.gelu_f32_(h)
        self.launches += 1
        ctx = torch.empty((B * T, self.D), dtype=f32, device=self.device)
        return self._linear(h, P["text_embedding.2.weight"], P["text_embedding.2.bias"], ctx, "ctx")

    # -- the step -----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, x, t, context, seq_len, y, full_ref, additional_control, density, block_hook=None,
                teacache=None, cond_flag=True) -> torch.Tensor:
        """Returns the stacked prediction [B, out_dim, F, H, W] in fp32. Latent/control inputs are taken as bf16
        tensors (the synthetic inputs are bf16-representable), everything downstream is fp32."""
        if teacache is not None or self.par is not None:
            raise FlexamNativeError("the fp32 verification engine runs on one GPU without TeaCache")
        cfg, D, dev = self.cfg, self.D, self.device
        self.refresh_if_modified()
        self.launches = 0
        C = cfg["out_dim"]
        x = x.to(dev, bf16).contiguous()
        y = y.to(dev, bf16).contiguous()
        add = additional_control.to(dev, bf16).contiguous()
        full_ref = full_ref.to(dev, bf16).contiguous()
        context = [u.to(dev, bf16) for u in context]
        t = t.to(dev, f32)
        density = density.to(dev, f32)
        B, _, F, Hh, Ww = x.shape
        Hp, Wp = Hh // 2, Ww // 2
        L0, R = F * Hp * Wp, Hp * Wp
        L = L0 + R
        if seq_len != L0:
            raise FlexamNativeError(f"seq_len {seq_len} != tokens on the grid {L0}")
        grid = (F + 1, Hp, Wp)
        M = B * L

        cnn = [self._cnn_fuser_f32(y[b, :C], add[b]) for b in range(B)]
        ctx = self._context_f32(context)
        T = cfg["text_len"]
        kvs = []
        for li, w in enumerate(self.blk):
            kv = torch.empty((B * T, 2 * D), dtype=f32, device=dev)
            self._linear(ctx, w["cwkv"], w["cbkv"], kv, "ctx")
            ops.rmsnorm_rope_f32(kv[:, :D], w["cnk"], self.eps)
            self.launches += 1
            kvs.append(kv)

        # patch + ref embedding (:885-899): three planes of [x | cnn | y_rest] patch rows, ref rows are one exact plane
        xs = self._buf("x", (M, D), f32)
        zx, zy = torch.zeros_like(x[0]), torch.zeros_like(y[0, C:])
        for b in range(B):
            o = b * L
            cplanes = self._split(cnn[b], "cnnout")
            rows = []
            for i in range(3):
                r = self._buf(f"patch_rows{i}", (L0, self.w_patch.shape[1]), bf16)
                ops.patchify([x[b] if i == 0 else zx, cplanes[i].view(F, Hh, Ww, -1), y[b, C:] if i == 0 else zy],
                             [False, True, False], F, Hh, Ww, r)
                rows.append(r)
            self._linear_planes(rows, self.w_patch, self.params["patch_embedding.bias"], xs[o + R: o + L])
            rrows = self._buf("ref_rows", (R, self.w_ref.shape[1]), bf16)
            ops.patchify([full_ref[b].unsqueeze(1)], [False], 1, Hh, Ww, rrows)
            self._linear_planes([rrows], self.w_ref, self.params["ref_conv.bias"], xs[o: o + R])
            self.launches += 4

        if t.dim() == 2:
            if t.shape[1] < L:
                t = torch.cat([t[:, -1:].expand(B, L - t.shape[1]), t], dim=1)
            uniq, inv = torch.unique(t.reshape(-1), return_inverse=True)
            row_idx = inv.to(i32).contiguous().view(-1)
        else:
            uniq = t.contiguous()
            row_idx = torch.arange(B, device=dev, dtype=i32).view(B, 1).expand(B, L).contiguous().view(-1)
        e, e0 = self._embed_mlp("time_embedding", "time_projection", uniq.contiguous())
        de, de0 = self._embed_mlp("density_embedding", "density_projection", density.contiguous())
        e0v, de0v = e0.view(-1, 6, D), de0.view(B, 2, D)

        h = self._buf("h32", (M, D), f32)
        qkv = self._buf("qkv32", (M, 3 * D), f32)
        attn = self._buf("attn32", (M, D), f32)
        cq = self._buf("cq32", (M, D), f32)
        yb = self._buf("y32", (M, D), f32)
        ffn = self._buf("ffn32", (M, cfg["ffn_dim"]), f32)
        scale = 1.0 / math.sqrt(128.0)
        qkv5 = qkv.view(B, L, 3, self.H, 128)
        attn4, cq4 = attn.view(B, L, self.H, 128), cq.view(B, L, self.H, 128)
        for li, w in enumerate(self.blk):
            mod, dmod = w["mod"], w["dmod"]
            ops.ln_f32(xs, h, self.eps, mod[0], mod[1], e0v[:, 0], e0v[:, 1], 6 * D, row_idx, dmod[0], de0v[:, 0],
                       2 * D, L)
            self._linear(h, w["wqkv"], w["bqkv"], qkv)
            ops.rmsnorm_rope_f32(qkv[:, :2 * D], w["nq"], self.eps, self.freqs, grid, 0, L, weight2=w["nk"])
            ops.attention_f32(qkv5[:, :, 0], qkv5[:, :, 1], qkv5[:, :, 2], attn4, scale)
            self._linear(attn, w["wo"], w["bo"], yb)
            ops.gated_residual_f32_(xs, yb, gate_mod=mod[2], gate_e=e0v[:, 2], row_idx=row_idx)
            ops.ln_f32(xs, h, self.eps, gamma=w["n3w"], beta=w["n3b"])
            self._linear(h, w["cwq"], w["cbq"], cq)
            ops.rmsnorm_rope_f32(cq, w["cnq"], self.eps)
            kv5 = kvs[li].view(B, T, 2, self.H, 128)
            ops.attention_f32(cq4, kv5[:, :, 0], kv5[:, :, 1], attn4, scale)
            self._linear(attn, w["cwo"], w["cbo"], yb)
            ops.gated_residual_f32_(xs, yb)
            ops.ln_f32(xs, h, self.eps, mod[3], mod[4], e0v[:, 3], e0v[:, 4], 6 * D, row_idx, dmod[1], de0v[:, 1],
                       2 * D, L)
            self._linear(h, w["w1"], w["b1"], ffn)
            ops.gelu_f32_(ffn)
            self._linear(ffn, w["w2"], w["b2"], yb, "ffn")
            ops.gated_residual_f32_(xs, yb, gate_mod=mod[5], gate_e=e0v[:, 5], row_idx=row_idx)
            self.launches += 11
            if block_hook is not None:
                block_hook(li, xs)

        ops.ln_f32(xs, h, self.eps, self.head_mod[0], self.head_mod[1], e, e, D, row_idx, self.head_dmod, de, D, L)
        ho = self._buf("head32", (M, self.params["head.head.weight"].shape[0]), f32)
        self._linear(h, self.params["head.head.weight"], self.params["head.head.bias"], ho)
        out = torch.empty((B, C, F, Hh, Ww), dtype=f32, device=dev)
        oplanes = torch.empty((3, C, F, Hh, Ww), dtype=bf16, device=dev)
        for b in range(B):
            hp = self._split(ho[b * L + R: (b + 1) * L], "head")
            for i in range(3):
                ops.unpatchify(hp[i], oplanes[i])
            ops.join3(oplanes, out[b])
            self.launches += 4
        return out


def precise_engine(model) -> PreciseEngine:
    """The fp32 verification engine over a model's parameters (a ``flexam_b200`` mirror or an ``install``-ed module)."""
    params = {k: v.detach() for k, v in model.named_parameters()}
    dev = next(iter(params.values())).device
    cfg = dict(model.config)
    cfg.setdefault("in_dim_ref_conv", params["ref_conv.weight"].shape[1])
    eng = PreciseEngine(params, cfg, dev)
    from .model import _sync_rope_table
    _sync_rope_table(model, eng)         # RoPE table from the module's `freqs` (RIFLEx on / off), like the bf16 path
    return eng
