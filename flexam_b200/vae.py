"""Wan2.2 VAE (decoder and encoder) on the native kernels.

``AutoencoderKLWan3_8`` mirrors the reference wrapper (FlexAM/models/wan_vae3_8.py:892-1057, cited as :line) for what the
pipeline calls — ``vae.encode(x).latent_dist`` and ``vae.decode(latents).sample`` — with the reference's parameter names
(``model.encoder.*``, ``model.conv1.*``, ``model.conv2.*``, ``model.decoder.*``). The forward is
``VaeDecoderEngine.decode``: the reference's frame-by-frame loop with its per-convolution feature cache (:820-849), every convolution an implicit GEMM on the tcgen05 kernels
(``fx_conv_gemm_bf16``: causal 3x3x3, per-frame 3x3, (3,1,1) time convolution; 1x1 as plain GEMMs) reading zero-padded
channel-last grids whose first two frames are the causal history, and the element-wise work between them (RMS_norm + SiLU,
nearest 2x, temporal interleave, DupUp3D shortcut, residual adds, the attention block's softmax, unpatchify + clamp) as
row kernels (csrc/vae.cu). ``encode`` (:788-819) is the mirror image: patchify, first frame alone then four frames per chunk
through Encoder3d (:564-618) — the same residual / attention blocks, the stride-2 convolutions of ``Resample`` downsample2d /
downsample3d as the stride forms of the implicit GEMM, AvgDown3D shortcuts — and ``conv1`` with the latent normalisation
folded in. With a ``flexam_b200.dist.SlabExchange`` attached the decode runs on one band of image rows per rank (halo rows
exchanged per 3x3 convolution). No torch compute, no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import ops
from .lib import FX_EPI_BF16, FX_EPI_F32_EXACT, FlexamNativeError

bf16, f32 = torch.bfloat16, torch.float32


def _pad64(c: int) -> int:
    return -(-c // 64) * 64


def decoder_dims(cfg: dict) -> List[int]:
    d, mult = cfg["dec_dim"], cfg["dim_mult"]
    return [d * u for u in [mult[-1]] + mult[::-1]]            # :642


def param_shapes(cfg: dict) -> Dict[str, tuple]:
    """state_dict keys / shapes of ``conv2`` + ``decoder`` of AutoencoderKLWan2_2_ (:739-782, :621-675)."""
    z, dims = cfg["z_dim"], decoder_dims(cfg)
    t_up = list(cfg["temperal_downsample"])[::-1]
    out: Dict[str, tuple] = {}

    def conv(name, co, ci, k):
        out[name + ".weight"] = (co, ci) + tuple(k)
        out[name + ".bias"] = (co,)

    def res(name, ci, co):
        out[name + ".residual.0.gamma"] = (ci, 1, 1, 1)
        conv(name + ".residual.2", co, ci, (3, 3, 3))
        out[name + ".residual.3.gamma"] = (co, 1, 1, 1)
        conv(name + ".residual.6", co, co, (3, 3, 3))
        if ci != co:
            conv(name + ".shortcut", co, ci, (1, 1, 1))

    conv("conv2", z, z, (1, 1, 1))
    conv("decoder.conv1", dims[0], z, (3, 3, 3))
    res("decoder.middle.0", dims[0], dims[0])
    out["decoder.middle.1.norm.gamma"] = (dims[0], 1, 1)
    conv("decoder.middle.1.to_qkv", 3 * dims[0], dims[0], (1, 1))
    conv("decoder.middle.1.proj", dims[0], dims[0], (1, 1))
    res("decoder.middle.2", dims[0], dims[0])
    n = len(cfg["dim_mult"])
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        for j in range(cfg["num_res_blocks"] + 1):
            res(f"decoder.upsamples.{i}.upsamples.{j}", ci if j == 0 else co, co)
        if i != n - 1:
            j = cfg["num_res_blocks"] + 1
            conv(f"decoder.upsamples.{i}.upsamples.{j}.resample.1", co, co, (3, 3))
            if i < len(t_up) and t_up[i]:
                conv(f"decoder.upsamples.{i}.upsamples.{j}.time_conv", 2 * co, co, (3, 1, 1))
    out["decoder.head.0.gamma"] = (dims[-1], 1, 1, 1)
    conv("decoder.head.2", 12, dims[-1], (3, 3, 3))
    return out


def encoder_dims(cfg: dict) -> List[int]:
    return [cfg["enc_dim"] * u for u in [1] + list(cfg["dim_mult"])]          # :527


def encoder_param_shapes(cfg: dict) -> Dict[str, tuple]:
    """state_dict keys / shapes of ``encoder`` + ``conv1`` of AutoencoderKLWan2_2_ (:505-562, :771)."""
    z, dims = cfg["z_dim"], encoder_dims(cfg)
    t_dn = list(cfg["temperal_downsample"])
    out: Dict[str, tuple] = {}

    def conv(name, co, ci, k):
        out[name + ".weight"] = (co, ci) + tuple(k)
        out[name + ".bias"] = (co,)

    def res(name, ci, co):
        out[name + ".residual.0.gamma"] = (ci, 1, 1, 1)
        conv(name + ".residual.2", co, ci, (3, 3, 3))
        out[name + ".residual.3.gamma"] = (co, 1, 1, 1)
        conv(name + ".residual.6", co, co, (3, 3, 3))
        if ci != co:
            conv(name + ".shortcut", co, ci, (1, 1, 1))

    conv("encoder.conv1", dims[0], 12, (3, 3, 3))
    n = len(cfg["dim_mult"])
    for i, (ci, co) in enumerate(zip(dims[:-1], dims[1:])):
        for j in range(cfg["num_res_blocks"]):
            res(f"encoder.downsamples.{i}.downsamples.{j}", ci if j == 0 else co, co)
        if i != n - 1:
            j = cfg["num_res_blocks"]
            conv(f"encoder.downsamples.{i}.downsamples.{j}.resample.1", co, co, (3, 3))
            if i < len(t_dn) and t_dn[i]:
                conv(f"encoder.downsamples.{i}.downsamples.{j}.time_conv", co, co, (3, 1, 1))
    res("encoder.middle.0", dims[-1], dims[-1])
    out["encoder.middle.1.norm.gamma"] = (dims[-1], 1, 1)
    conv("encoder.middle.1.to_qkv", 3 * dims[-1], dims[-1], (1, 1))
    conv("encoder.middle.1.proj", dims[-1], dims[-1], (1, 1))
    res("encoder.middle.2", dims[-1], dims[-1])
    out["encoder.head.0.gamma"] = (dims[-1], 1, 1, 1)
    conv("encoder.head.2", 2 * z, dims[-1], (3, 3, 3))
    conv("conv1", 2 * z, 2 * z, (1, 1, 1))
    return out


class VaeDecoderEngine:
    """Packed weights, the per-convolution history grids and the launch sequences of decode and encode."""

    def __init__(self, params: Dict[str, torch.Tensor], cfg: dict, device: torch.device):
        self.cfg, self.device, self.params = dict(cfg), torch.device(device), params
        self.dims = decoder_dims(cfg)
        for c in self.dims + (encoder_dims(cfg) if "enc_dim" in cfg else []):
            if c % 8 != 0:
                raise FlexamNativeError(f"native VAE: channel widths must be multiples of 8, got {c}")
        self._ws: Dict[tuple, torch.Tensor] = {}
        self._hist: Dict[str, torch.Tensor] = {}
        self._geom: Dict[str, tuple] = {}
        self.launches = 0
        self.slab = None            # flexam_b200.dist.SlabExchange: decode split into bands of image rows across the ranks
        self._slab_on = False
        self._pack()

    # -- weights: tap-major [Cout (padded to 8), taps * Cin (padded to 64)] ------------------------------------------
    def _pack(self):
        P, dev = self.params, self.device
        for k, v in P.items():
            if v.device != dev or v.dtype != bf16:
                raise FlexamNativeError(f"parameter {k}: expected bf16 on {dev}, got {v.dtype} on {v.device}")
        self.w: Dict[str, torch.Tensor] = {}
        self.b: Dict[str, torch.Tensor] = {}
        for k, v in P.items():
            if not k.endswith(".weight"):
                continue
            name = k[:-7]
            co, ci = v.shape[:2]
            taps = int(math.prod(v.shape[2:]))
            # k-blocks of the implicit GEMM are 64 channels of ONE tap: pad Cin; 1x1 convolutions are plain GEMMs (K = Cin)
            cip, cop = (ci if taps == 1 else _pad64(ci)), -(-co // 8) * 8
            wt = torch.zeros((cop, taps, cip), dtype=bf16, device=dev)
            wt[:co, :, :ci] = v.reshape(co, ci, taps).permute(0, 2, 1)          # K order (dt, dy, dx, cin)
            self.w[name] = wt.view(cop, taps * cip).contiguous()
            bias = torch.zeros((cop,), dtype=bf16, device=dev)
            bias[:co] = P[name + ".bias"]
            self.b[name] = bias
        self.gamma = {k[:-6]: v.reshape(-1).contiguous() for k, v in P.items() if k.endswith(".gamma")}
        # attention blocks: q|k and v as separate projections (v must be contiguous for the transpose)
        self.attn_w = {}
        for k, v in P.items():
            if k.endswith(".to_qkv.weight"):
                name = k[:-len(".to_qkv.weight")]
                C = v.shape[1]
                wq, bq = v.reshape(3 * C, C), P[name + ".to_qkv.bias"]
                self.attn_w[name] = (wq[:2 * C].contiguous(), bq[:2 * C].contiguous(), wq[2 * C:].contiguous(),
                                     bq[2 * C:].contiguous())
        self._versions = tuple(p._version for p in P.values())
        self._scale_key = None
        self._enc_scale_key = None

    def _fold_scale(self, scale: Sequence[torch.Tensor]):
        """conv2(z / scale[1] + scale[0]) (:824-831) as ONE projection of the raw latents: the per-channel affine is
        folded into conv2's weight and bias in fp32 once per (mean, 1/std) pair (weight-side constant folding)."""
        key = (scale[0].data_ptr(), scale[1].data_ptr(), scale[0]._version, scale[1]._version)
        if self._scale_key == key:
            return
        zd = self.cfg["z_dim"]
        w = self.params["conv2.weight"].reshape(zd, zd).float()
        mean, inv_std = scale[0].to(self.device, f32), scale[1].to(self.device, f32)
        wf = torch.zeros((_pad64(zd), _pad64(zd)), dtype=bf16, device=self.device)
        wf[:zd, :zd] = (w / inv_std.view(1, zd)).to(bf16)
        bfold = torch.zeros((_pad64(zd),), dtype=bf16, device=self.device)
        bfold[:zd] = (self.params["conv2.bias"].float() + w @ mean).to(bf16)
        self.w_conv2, self.b_conv2 = wf, bfold
        self._scale_key = key
        self._scale_ref = (scale[0], scale[1])

    # -- buffers ---------------------------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype=bf16, zero=False):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    def _grid(self, name: str, frames: int, Hp: int, Wp: int, C: int, keep_rows: int = 0) -> torch.Tensor:
        """The zero-padded input grid of convolution ``name``: [frames, Hp, Wp, C] flattened to rows. Halo positions and
        the history frames start as zeros and only interior / live positions are ever written. A chunk with more frames
        than any before (the first chunk is not up-sampled in time) grows the grid and keeps its ``keep_rows`` history."""
        need = frames * Hp * Wp
        key = ("slab:" if self._slab_on else "") + name       # band-sized, peer-mapped grids live next to the full-size ones
        t = self._hist.get(key)
        same_plane = self._geom.get(key) == (Hp, Wp)
        self._geom[key] = (Hp, Wp)
        if t is None or t.shape[0] < need or t.shape[1] != C:
            if self._slab_on:
                new = self.slab.alloc(key, need, C, self.device)       # zeroed; collective (every rank grows the same grid)
            else:
                new = torch.zeros((need, C), dtype=bf16, device=self.device)
            # the history travels only when the grid grows inside one clip (same plane geometry); a grid left over from a
            # clip of another size starts from zeros (decode / encode cleared it anyway)
            if t is not None and t.shape[1] == C and keep_rows and same_plane and t.shape[0] >= keep_rows:
                new[:keep_rows].copy_(t[:keep_rows])
            self._hist[key] = t = new
        return t

    def _reset_history(self):
        for t in self._hist.values():
            t.zero_()
        if self._slab_on:
            self.slab.sync()                 # nobody pushes a halo row into a grid its owner has not cleared yet

    def _halo(self, grid, frame0, T, Hp, Wp):
        """Slab decode: the first / last interior row of my band is the halo row of the band above / below."""
        if self._slab_on:
            self.slab.halo(grid, frame0, T, Hp, Wp)
            self.launches += 2

    # -- building blocks -------------------------------------------------------------------------------------------
    def _cconv(self, name, x, T, H, W, gamma=None, silu=False, kt=3, ks=3, stride_s=1, stride_t=1, run=True):
        """[RMS_norm + SiLU ->] causal convolution ``name`` over the chunk x (dense [T*H*W, Cin]) with the two cached
        history frames in front (CausalConv3d :22-47 + the cache rule of its callers :219-238): returns dense
        [T*H*W, Cout]. The history grid then keeps the last two input frames for the next chunk."""
        w = self.w[name]
        cin = w.shape[1] // (kt * ks * ks)
        pad = 1 if ks == 3 else 0
        Hp, Wp = H + 2 * pad, W + 2 * pad
        hist = kt - 1
        grid = self._grid(name, hist + T, Hp, Wp, cin, keep_rows=hist * Hp * Wp)
        ops.vae_norm_act(x, gamma, grid, H, W, pad, hist, silu)
        if pad:
            self._halo(grid, hist, T, Hp, Wp)
        out = None
        if run:
            out = self._buf("out:" + name, ((T // stride_t) * (H // stride_s) * (W // stride_s), w.shape[0]))
            rows = (hist + T) * Hp * Wp
            ops.conv_gemm(grid[:rows], w, self.b[name], out, T, H, W, kt, ks, FX_EPI_BF16, stride_s, stride_t)
        self.launches += 2 if run else 1
        if hist:
            plane = Hp * Wp
            src = grid[T * plane:(T + hist) * plane]
            grid[:hist * plane].copy_(src.clone() if T < hist else src)       # last two frames become the history
        return out

    def _res(self, name, x, T, H, W):
        """ResidualBlock (:198-240)."""
        if name + ".shortcut" in self.w:
            w = self.w[name + ".shortcut"]
            h = self._buf("sc:" + name, (x.shape[0], w.shape[0]))
            ops.gemm(x, w, self.b[name + ".shortcut"], h, FX_EPI_BF16)
            self.launches += 1
        else:
            h = x
        y = self._cconv(name + ".residual.2", x, T, H, W, self.gamma[name + ".residual.0"], True)
        y = self._cconv(name + ".residual.6", y, T, H, W, self.gamma[name + ".residual.3"], True)
        ops.add_bf16_(y, h)
        self.launches += 1
        return y

    def _attn(self, name, x, T, H, W):
        """AttentionBlock (:243-282): one head of width C over the H*W tokens of each frame. S = Q K^T and O = P V are
        GEMMs on the tensor cores (V transposed once into the [C, tokens] weight layout), softmax a row kernel."""
        C, P = x.shape[1], H * W
        if self._slab_on:        # every rank attends over the whole frame (25 GFLOP): gather the bands, keep my rows
            ex = self.slab
            if T != 1:
                raise FlexamNativeError("internal: slab decode expects one frame per chunk at the attention block")
            full = ex.gather_rows(x.contiguous(), self._buf("attn_gather", (ex.world * P, C)))
            self._slab_on = False
            try:
                y = self._attn(name, full, 1, H * ex.world, W)
            finally:
                self._slab_on = True
            self.launches += 1
            return y[ex.rank * P:(ex.rank + 1) * P]
        out = self._buf("attn_out", (T * P, C))
        # the score / probability matrices are GEMM operands: their key dimension is padded to a multiple of 8 (zero K and
        # V rows, zero probabilities; the softmax runs over the P real keys only). No effect when H * W % 8 == 0.
        Pp = -(-P // 8) * 8
        pad = Pp != P
        for f in range(T):
            xf = x[f * P:(f + 1) * P]
            y = self._buf("attn_norm", (P, C))
            ops.vae_norm_act(xf, self.gamma[name + ".norm"], y, H, W, 0, 0, False)
            qk = self._buf("attn_qk", (Pp, 2 * C), zero=pad)
            v = self._buf("attn_v", (Pp, C), zero=pad)
            w_qk, b_qk, w_v, b_v = self.attn_w[name]
            ops.gemm(y, w_qk, b_qk, qk[:P], FX_EPI_BF16)
            ops.gemm(y, w_v, b_v, v[:P], FX_EPI_BF16)
            s = self._buf("attn_s", (P, Pp), f32)
            ops.gemm(qk[:P, :C], qk[:, C:], None, s, FX_EPI_F32_EXACT)
            p = self._buf("attn_p", (P, Pp), zero=pad)
            ops.softmax_rows(s[:, :P], p[:, :P], 1.0 / math.sqrt(C))
            vt = self._buf("attn_vt", (C, Pp))
            ops.nchw_to_nhwc(v, vt, 0)                                    # [Pp, C] -> [C, Pp]
            o = self._buf("attn_o", (P, C))
            ops.gemm(p, vt, None, o, FX_EPI_BF16)
            of = out[f * P:(f + 1) * P]
            ops.gemm(o, self.w[name + ".proj"], self.b[name + ".proj"], of, FX_EPI_BF16)
            ops.add_bf16_(of, xf.contiguous())
            self.launches += 9
        return out

    def _resample(self, name, x, T, H, W, temporal, first):
        """Resample upsample2d / upsample3d (:117-160)."""
        C = x.shape[1]
        if temporal and not first:        # the first chunk is not up-sampled in time (the "Rep" rule :121-123)
            y = self._cconv(name + ".time_conv", x, T, H, W, None, False, kt=3, ks=1)            # [T*HW, 2C]
            x2 = self._buf("tint:" + name, (2 * T * H * W, C))
            ops.vae_time_interleave(y, x2, T, H * W)
            self.launches += 1
            x, T = x2, 2 * T
        up = self._grid(name + ".resample.1", T, 2 * H + 2, 2 * W + 2, C)
        rows = T * (2 * H + 2) * (2 * W + 2)
        ops.vae_upsample2x(x.contiguous(), up[:rows], T, H, W)
        self._halo(up, 0, T, 2 * H + 2, 2 * W + 2)
        w = self.w[name + ".resample.1"]
        out = self._buf("out:" + name, (T * 4 * H * W, w.shape[0]))
        ops.conv_gemm(up[:rows], w, self.b[name + ".resample.1"], out, T, 2 * H, 2 * W, 1, 3, FX_EPI_BF16)
        self.launches += 2
        return out, T

    # -- decode ----------------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: torch.Tensor, scale: Sequence[torch.Tensor]) -> torch.Tensor:
        """z: [1, z_dim, T, H, W] normalised latents -> bf16 video [1, 3, 1 + 4 (T - 1), 16 H, 16 W] in [-1, 1]."""
        if tuple(p._version for p in self.params.values()) != self._versions:
            self._pack()
        cfg, dev = self.cfg, self.device
        zd = cfg["z_dim"]
        if z.dim() != 5 or z.shape[0] != 1 or z.shape[1] != zd:
            raise FlexamNativeError(f"VAE decode: expected latents [1, {zd}, T, H, W], got {tuple(z.shape)}")
        self._fold_scale(scale)
        _, _, T, H, W = z.shape
        ex = self.slab
        self._slab_on = ex is not None and ex.world > 1 and H % ex.world == 0
        try:
            return self._decode(z, T, H // ex.world if self._slab_on else H, W)
        finally:
            self._slab_on = False

    def _decode(self, z, T, H, W):
        """H: the rows of the latent grid this rank decodes (all of them, or its band of a slab decode)."""
        cfg, dev, zd, ex = self.cfg, self.device, self.cfg["z_dim"], self.slab
        self._reset_history()
        self.launches = 0
        dims = self.dims
        n = len(cfg["dim_mult"])
        t_up = list(cfg["temperal_downsample"])[::-1]
        nres = cfg["num_res_blocks"] + 1
        P = H * W
        zb = z[0, :, :, ex.rank * H:(ex.rank + 1) * H] if self._slab_on else z[0]
        zc = zb.to(dev, bf16).contiguous().view(zd, T * P)
        zl = self._buf("z_rows", (T * P, _pad64(zd)), zero=True)
        ops.nchw_to_nhwc(zc, zl, 0)
        x_all = self._buf("conv2_out", (T * P, _pad64(zd)))
        ops.gemm(zl, self.w_conv2, self.b_conv2, x_all, FX_EPI_BF16)            # un-normalise + conv2 (:824-831)
        self.launches += 2
        Tout = 1 + 4 * (T - 1)
        up_s = 2 ** (n - 1)
        video = torch.empty((3, Tout, 2 * up_s * H, 2 * up_s * W), dtype=bf16, device=dev)
        f_out = 0
        for i in range(T):                                                       # :832-846, one latent frame per chunk
            first = i == 0
            Tc, h, w = 1, H, W
            x = self._cconv("decoder.conv1", x_all[i * P:(i + 1) * P], Tc, h, w)
            x = self._res("decoder.middle.0", x, Tc, h, w)
            x = self._attn("decoder.middle.1", x, Tc, h, w)
            x = self._res("decoder.middle.2", x, Tc, h, w)
            for s in range(n):                                                   # Up_ResidualBlock :494-502
                name = f"decoder.upsamples.{s}.upsamples."
                up = s != n - 1
                temporal = up and s < len(t_up) and bool(t_up[s])
                x_in = x
                main = x
                for j in range(nres):
                    main = self._res(name + str(j), main, Tc, h, w)
                if up:
                    if main is x_in or main.data_ptr() == x_in.data_ptr():
                        raise FlexamNativeError("internal: residual output aliases the block input")
                    main, T2 = self._resample(name + str(nres), main, Tc, h, w, temporal, first)
                    ft = 2 if temporal else 1
                    ops.vae_dupup_add_(main, x_in.contiguous(), T2, h, w, ft, ft - 1 if first else 0)
                    self.launches += 1
                    Tc, h, w = T2, 2 * h, 2 * w
                x = main
            y = self._cconv("decoder.head.2", x, Tc, h, w, self.gamma["decoder.head.0"], True)
            ops.vae_unpatchify(y, video, Tc, h, w, f_out)
            self.launches += 1
            f_out += Tc
        if f_out != Tout:
            raise FlexamNativeError(f"internal: decoded {f_out} frames, expected {Tout}")
        if self._slab_on:
            video = ex.gather_video(video)                # [3, Tout, world * band rows, width] on every rank
        return video.unsqueeze(0)


    # -- encode ----------------------------------------------------------------------------------------------------
    def _fold_enc_scale(self, scale: Sequence[torch.Tensor]):
        """conv1 followed by (mu - mean) * (1 / std) on the first z_dim channels (:811-816) as ONE projection: the affine
        is folded into conv1's weight rows and bias in fp32 once per (mean, 1/std) pair."""
        key = (scale[0].data_ptr(), scale[1].data_ptr(), scale[0]._version, scale[1]._version)
        if self._enc_scale_key == key:
            return
        zd = self.cfg["z_dim"]
        w = self.params["conv1.weight"].reshape(2 * zd, 2 * zd).float().clone()
        b = self.params["conv1.bias"].float().clone()
        mean, inv_std = scale[0].to(self.device, f32), scale[1].to(self.device, f32)
        w[:zd] *= inv_std.view(zd, 1)
        b[:zd] = (b[:zd] - mean) * inv_std
        self.w_conv1, self.b_conv1 = w.to(bf16).contiguous(), b.to(bf16).contiguous()
        self._enc_scale_key = key
        self._enc_scale_ref = (scale[0], scale[1])

    def _downsample(self, name, x, T, H, W, temporal, first):
        """Resample downsample2d / downsample3d (:117-160): ZeroPad2d((0,1,0,1)) + 3x3 stride-2 convolution per frame,
        then (3,1,1) stride-(2,1,1) convolution over [last cached frame | chunk] (the first chunk only seeds the cache)."""
        C = x.shape[1]
        grid = self._grid(name + ".resample.1", T, H + 2, W + 2, _pad64(C))
        rows = T * (H + 2) * (W + 2)
        ops.vae_norm_act(x, None, grid, H, W, 1, 0, False)
        w = self.w[name + ".resample.1"]
        out = self._buf("out:" + name, (T * (H // 2) * (W // 2), w.shape[0]))
        ops.conv_gemm(grid[:rows], w, self.b[name + ".resample.1"], out, T, H, W, 1, 3, FX_EPI_BF16, 2, 1)
        self.launches += 2
        H, W = H // 2, W // 2
        if temporal:
            if first:
                self._cconv(name + ".time_conv", out, T, H, W, None, False, kt=3, ks=1, run=False)     # :147-149
            else:
                out = self._cconv(name + ".time_conv", out, T, H, W, None, False, kt=3, ks=1, stride_t=2)
                T //= 2
        return out, T, H, W

    @torch.no_grad()
    def encode(self, x: torch.Tensor, scale: Sequence[torch.Tensor]) -> torch.Tensor:
        """x: [1, 3, T, H, W] pixels in [-1, 1] (T = 1 + 4k, H and W multiples of 16) -> bf16 [1, 2 z_dim, 1 + k, H/16,
        W/16]: normalised mean | log-variance (:788-819)."""
        if "enc_dim" not in self.cfg or "encoder.conv1.weight" not in self.params:
            raise FlexamNativeError("this VAE engine was built without encoder parameters")
        if tuple(p._version for p in self.params.values()) != self._versions:
            self._pack()
        cfg, dev = self.cfg, self.device
        zd = cfg["z_dim"]
        n = len(cfg["dim_mult"])
        if x.dim() != 5 or x.shape[0] != 1 or x.shape[1] != 3 or (x.shape[2] - 1) % 4 != 0 or \
                x.shape[3] % (2 ** n) != 0 or x.shape[4] % (2 ** n) != 0:
            raise FlexamNativeError(f"VAE encode: expected video [1, 3, 1+4k, 16a, 16b], got {tuple(x.shape)}")
        self._fold_enc_scale(scale)
        self._reset_history()
        self.launches = 0
        _, _, T, H, W = x.shape
        dims = encoder_dims(cfg)
        t_dn = list(cfg["temperal_downsample"])
        nres = cfg["num_res_blocks"]
        vid = x[0].to(dev, bf16).contiguous()
        h0, w0 = H // 2, W // 2
        chunks = 1 + (T - 1) // 4
        hl, wl = H // (2 ** n), W // (2 ** n)
        lat = self._buf("enc_lat", (chunks * hl * wl, 2 * zd))
        for i in range(chunks):                                                  # :795-810
            first = i == 0
            f0, Tc = (0, 1) if first else (1 + 4 * (i - 1), 4)
            h, w = h0, w0
            rows = self._buf("enc_rows", (Tc * h * w, 64), zero=True)            # 12 patch channels, zero-padded to 64
            ops.vae_patchify(vid, rows, Tc, h, w, f0)
            self.launches += 1
            y = self._cconv("encoder.conv1", rows, Tc, h, w)
            for s in range(n):                                                   # Down_ResidualBlock :452-457
                name = f"encoder.downsamples.{s}.downsamples."
                down = s != n - 1
                temporal = down and s < len(t_dn) and bool(t_dn[s])
                y_in, T_in, h_in, w_in = y, Tc, h, w
                for j in range(nres):
                    y = self._res(name + str(j), y, Tc, h, w)
                if down:
                    y, Tc, h, w = self._downsample(name + str(nres), y, Tc, h, w, temporal, first)
                ops.vae_avgdown_add_(y, y_in.contiguous(), T_in, h_in, w_in, 2 if temporal else 1, 2 if down else 1)
                self.launches += 1
            y = self._res("encoder.middle.0", y, Tc, h, w)
            y = self._attn("encoder.middle.1", y, Tc, h, w)
            y = self._res("encoder.middle.2", y, Tc, h, w)
            y = self._cconv("encoder.head.2", y, Tc, h, w, self.gamma["encoder.head.0"], True)
            if Tc != 1 or (h, w) != (hl, wl):
                raise FlexamNativeError(f"internal: encoder chunk ended at {Tc} x {h} x {w}")
            lat[i * hl * wl:(i + 1) * hl * wl].copy_(y[:, :2 * zd])
        out_rows = self._buf("enc_out_rows", (chunks * hl * wl, 2 * zd))
        ops.gemm(lat, self.w_conv1, self.b_conv1, out_rows, FX_EPI_BF16)         # conv1 + latent normalisation :811-816
        out = torch.empty((2 * zd, chunks * hl * wl), dtype=bf16, device=dev)
        ops.nchw_to_nhwc(out_rows, out, 0)                                       # [P, 2z] -> [2z, P] (channel-first)
        self.launches += 2
        return out.view(1, 2 * zd, chunks, hl, wl)


# ----------------------------------------------------------------------------------------------------------
# nn.Module with the reference wrapper's surface
# ----------------------------------------------------------------------------------------------------------
class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class AutoencoderKLOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class DiagonalGaussianDistribution:
    """The part of diffusers' class the FlexAM pipeline uses on ``vae.encode(x)[0]`` (``.mode()`` / ``.sample()``):
    parameters = mean | log-variance along dim 1, log-variance clamped to [-30, 20]."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


def _set_param(root: nn.Module, dotted: str, p: nn.Parameter):
    parts = dotted.split(".")
    m = root
    for name in parts[:-1]:
        child = m._modules.get(name)
        if child is None:
            child = nn.Module()
            m.add_module(name, child)
        m = child
    m.register_parameter(parts[-1], p)


# Per-channel statistics of the Wan2.2 latent space: model constants the reference wrapper carries (:906-1008)
WAN22_LATENTS_MEAN = [
    -0.2289, -0.0052, -0.1323, -0.2339, -0.2799, 0.0174, 0.1838, 0.1557, -0.1382, 0.0542, 0.2813, 0.0891, 0.1570, -0.0098,
    0.0375, -0.1825, -0.2246, -0.1207, -0.0698, 0.5109, 0.2665, -0.2108, -0.2158, 0.2502, -0.2055, -0.0322, 0.1109, 0.1567,
    -0.0729, 0.0899, -0.2799, -0.1230, -0.0313, -0.1649, 0.0117, 0.0723, -0.2839, -0.2083, -0.0520, 0.3748, 0.0152, 0.1957,
    0.1433, -0.2944, 0.3573, -0.0548, -0.1681, -0.0667]
WAN22_LATENTS_STD = [
    0.4765, 1.0364, 0.4514, 1.1677, 0.5313, 0.4990, 0.4818, 0.5013, 0.8158, 1.0344, 0.5894, 1.0901, 0.6885, 0.6165, 0.8454,
    0.4978, 0.5759, 0.3523, 0.7135, 0.6804, 0.5833, 1.4146, 0.8986, 0.5659, 0.7069, 0.5338, 0.4889, 0.4917, 0.4069, 0.4999,
    0.6866, 0.4093, 0.5709, 0.6065, 0.6415, 0.4944, 0.5726, 1.2042, 0.5458, 1.6887, 0.3971, 1.0600, 0.3943, 0.5537, 0.5444,
    0.4089, 0.7468, 0.7744]


class AutoencoderKLWan3_8(nn.Module):
    """Drop-in for FlexAM.models.AutoencoderKLWan3_8 (:892-1057) as the pipeline uses it: ``encode(x).latent_dist``
    (``.mode()`` / ``.sample()``) and ``decode(z).sample``. Parameters carry the reference's names under ``model.``
    (``model.encoder.*``, ``model.conv1.*``, ``model.conv2.*``, ``model.decoder.*``): the reference checkpoint loads
    strictly. ``enable_multi_gpus_inference()`` splits the decode over the ranks of a box (not in the reference)."""

    def __init__(self, latent_channels=48, c_dim=160, vae_pth=None, dim_mult=(1, 2, 4, 4),
                 temperal_downsample=(False, True, True), temporal_compression_ratio=4, spatial_compression_ratio=8,
                 dec_dim=256, latents_mean=None, latents_std=None, dtype=torch.bfloat16, device=None):
        super().__init__()
        self.cfg = dict(z_dim=latent_channels, dec_dim=dec_dim, enc_dim=c_dim, dim_mult=list(dim_mult), num_res_blocks=2,
                        temperal_downsample=list(temperal_downsample))
        self.temporal_compression_ratio = temporal_compression_ratio
        self.spatial_compression_ratio = spatial_compression_ratio
        for name, shape in {**encoder_param_shapes(self.cfg), **param_shapes(self.cfg)}.items():
            _set_param(self, "model." + name, nn.Parameter(torch.empty(shape, dtype=dtype, device=device),
                                                           requires_grad=False))
        if latents_mean is None and latents_std is None and latent_channels == len(WAN22_LATENTS_MEAN):
            latents_mean, latents_std = WAN22_LATENTS_MEAN, WAN22_LATENTS_STD      # the reference's constants (:906-1008)
        mean = torch.zeros(latent_channels) if latents_mean is None else torch.as_tensor(latents_mean, dtype=f32)
        std = torch.ones(latent_channels) if latents_std is None else torch.as_tensor(latents_std, dtype=f32)
        self.scale = [mean, 1.0 / std]
        self._engine: Optional[VaeDecoderEngine] = None
        self._slab = None

    def enable_multi_gpus_inference(self, group=None):
        """Split ``decode`` into bands of image rows over the ranks (``flexam_b200.dist.enable_vae_slabs``); the counterpart
        of the transformer's ``enable_multi_gpus_inference`` — the reference decodes the whole clip on every rank."""
        from .dist import enable_vae_slabs
        return enable_vae_slabs(self, group=group)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs={}):
        """The reference loader's contract (:1058-1079): constructor kwargs filtered from ``additional_kwargs``, a
        ``.safetensors`` or ``torch.load`` checkpoint of the inner model whose keys get the ``model.`` prefix, non-strict
        load with the missing / unexpected keys printed. (bf16 parameters: the pipeline casts the VAE to its weight dtype.)"""
        import inspect
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in dict(additional_kwargs).items() if k in valid})
        if str(pretrained_model_path).endswith(".safetensors"):
            from safetensors.torch import load_file
            state_dict = load_file(pretrained_model_path)
        else:
            state_dict = torch.load(pretrained_model_path, map_location="cpu")
        state_dict = {"model." + k: v for k, v in state_dict.items()}
        own = model.state_dict()
        m, u = model.load_state_dict({k: v.to(own[k].dtype) if k in own else v for k, v in state_dict.items()}, strict=False)
        print(f"### missing keys: {len(m)}; \n### unexpected keys: {len(u)};")
        print(m, u)
        return model

    def engine(self) -> VaeDecoderEngine:
        params = {k[len("model."):]: v.detach() for k, v in self.named_parameters()}
        if self._engine is None or tuple(p.data_ptr() for p in self._engine.params.values()) != \
                tuple(p.data_ptr() for p in params.values()):
            self._engine = VaeDecoderEngine(params, self.cfg, next(iter(params.values())).device)
        self._engine.slab = self._slab
        return self._engine

    def _decode(self, zs: torch.Tensor) -> DecoderOutput:                       # :1041-1049
        eng = self.engine()
        with (torch.cuda.device(eng.device) if eng.device.type == "cuda" else _null()), ops.stream_scope():
            dec = [eng.decode(u.unsqueeze(0), self.scale)[0] for u in zs]
        return DecoderOutput(torch.stack(dec))

    def decode(self, z: torch.Tensor, return_dict: bool = True):                # :1051-1057
        decoded = self._decode(z).sample
        return DecoderOutput(decoded) if return_dict else (decoded,)

    def _encode(self, x: torch.Tensor) -> torch.Tensor:                         # :1021-1028
        eng = self.engine()
        with (torch.cuda.device(eng.device) if eng.device.type == "cuda" else _null()), ops.stream_scope():
            return torch.stack([eng.encode(u.unsqueeze(0), self.scale)[0].clone() for u in x])

    def encode(self, x: torch.Tensor, return_dict: bool = True):                # :1030-1039
        posterior = DiagonalGaussianDistribution(self._encode(x))
        return AutoencoderKLOutput(posterior) if return_dict else (posterior,)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
