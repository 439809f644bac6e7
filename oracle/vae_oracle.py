"""TEST INFRASTRUCTURE — CPU restatement of the reference Wan2.2 VAE DECODER (decode half).

Restates ``AutoencoderKLWan2_2_.decode`` and everything under it (reference file ``FlexAM/models/wan_vae3_8.py``, cited
as :line): latent un-normalisation and ``conv2`` :820-830, the frame-by-frame loop with the per-convolution feature cache
:831-849, ``Decoder3d.forward`` :677-728, ``ResidualBlock`` :198-240, ``AttentionBlock`` :243-282, ``Resample``
(upsample2d / upsample3d with its "Rep" first-chunk rule) :76-160, ``Up_ResidualBlock`` + ``DupUp3D`` :375-502,
``CausalConv3d`` :22-47, ``RMS_norm`` :50-64, ``unpatchify`` :304-318 and the wrapper's clamp :1041-1049 — as one
functional pass over a ``{state_dict key: tensor}`` mapping. Only ``tests/`` and ``bench.py`` may import it.

Pinning: ``oracle/make_golden.py vae_tiny`` runs the REAL module (``oracle/ref_import.build_reference_vae``) on
``oracle/synth`` weights / latents and stores its fp32 CPU output; ``tests/test_vae.py`` checks this file against it and,
when /root/reference is mounted, against the live module. The ENCODER half of the row is not restated (not built).

Policies: "fp32", and "bf16" — the module as the pipeline runs it (bf16 weights and activations): results of the
convolutions, the norm / SiLU chains, the attention and the residual adds are rounded to bf16. Called with bf16 TENSORS
(policy "fp32": no extra rounding) the same code is the module's own bf16 execution with stock torch ops (cuDNN
convolutions, SDPA): bench.py's library-path leg for the decoder.
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

CACHE_T = 2


def _r(t, policy):
    return t.to(torch.bfloat16).to(torch.float32) if policy == "bf16" else t


def decoder_dims(cfg: dict) -> List[int]:
    d, mult = cfg["dec_dim"], cfg["dim_mult"]
    return [d * u for u in [mult[-1]] + mult[::-1]]            # :642


def param_specs(cfg: dict):
    """[(state_dict key, shape, std, mean)] of the decoder half of AutoencoderKLWan2_2_ (conv2 + decoder)."""
    z, dims = cfg["z_dim"], decoder_dims(cfg)
    t_up = cfg["temperal_downsample"][::-1]
    specs = []

    def conv(name, co, ci, k):
        fan = ci * math.prod(k)
        specs.append((name + ".weight", (co, ci) + tuple(k), fan ** -0.5, 0.0))
        specs.append((name + ".bias", (co,), 0.02, 0.0))

    def res(name, ci, co):
        specs.append((name + ".residual.0.gamma", (ci, 1, 1, 1), 0.1, 1.0))
        conv(name + ".residual.2", co, ci, (3, 3, 3))
        specs.append((name + ".residual.3.gamma", (co, 1, 1, 1), 0.1, 1.0))
        conv(name + ".residual.6", co, co, (3, 3, 3))
        if ci != co:
            conv(name + ".shortcut", co, ci, (1, 1, 1))

    conv("conv2", z, z, (1, 1, 1))
    conv("decoder.conv1", dims[0], z, (3, 3, 3))
    res("decoder.middle.0", dims[0], dims[0])
    specs.append(("decoder.middle.1.norm.gamma", (dims[0], 1, 1), 0.1, 1.0))
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
    specs.append(("decoder.head.0.gamma", (dims[-1], 1, 1, 1), 0.1, 1.0))
    conv("decoder.head.2", 12, dims[-1], (3, 3, 3))
    return specs


def state_dict(cfg: dict, tag: str = "vae"):
    from oracle import synth
    return {name: synth.tensor(f"{tag}/{name}", shape, std, mean) for name, shape, std, mean in param_specs(cfg)}


def state_dict_torch(cfg: dict, device, dtype=None, tag: str = "vae"):
    from oracle import synth
    return {name: synth.tensor_torch(f"{tag}/{name}", shape, std, mean, device=device, dtype=dtype)
            for name, shape, std, mean in param_specs(cfg)}


def latents(cfg: dict, T: int, H: int, W: int, tag: str = "vaez"):
    from oracle import synth
    return synth.tensor(f"{tag}/z", (1, cfg["z_dim"], T, H, W))


def latent_scale(cfg: dict):
    """(mean, 1/std) per latent channel: synthetic stand-ins for the constants of AutoencoderKLWan3_8 :906-1008."""
    from oracle import synth
    z = cfg["z_dim"]
    mean = torch.from_numpy(synth.tensor("vae/latent_mean", (z,), 0.2, 0.0, bf16=False))
    std = torch.from_numpy(synth.tensor("vae/latent_std", (z,), 0.15, 0.7, bf16=False))
    return [mean, 1.0 / std]


VAE_CONFIGS = {
    # Wan2.2 VAE as AutoencoderKLWan3_8 builds it (:1009-1017): z 48, decoder width 256, 4 x 16 x 16 compression
    "real": dict(z_dim=48, dec_dim=256, enc_dim=160, dim_mult=[1, 2, 4, 4], num_res_blocks=2,
                 temperal_downsample=[False, True, True]),
    # same topology, decoder width 64 (channels 256/256/256/128/64: still multiples of 64): CPU-runnable in seconds
    "tiny": dict(z_dim=48, dec_dim=64, enc_dim=32, dim_mult=[1, 2, 4, 4], num_res_blocks=2,
                 temperal_downsample=[False, True, True]),
}


class _Run:
    def __init__(self, sd, policy):
        self.sd, self.policy = sd, policy
        self.cache: Dict[str, object] = {}

    # CausalConv3d :22-47 with the cache rule of its callers (:219-238, :677-728)
    def cconv(self, name, x, cached=True):
        w, b = self.sd[name + ".weight"], self.sd[name + ".bias"]
        kt, kh, kw = w.shape[2:]
        pad_t = kt - 1
        if cached and pad_t > 0:
            prev = self.cache.get(name)
            cache_x = x[:, :, -CACHE_T:].clone()
            if cache_x.shape[2] < 2 and prev is not None:
                cache_x = torch.cat([prev[:, :, -1:], cache_x], dim=2)
            xin = x if prev is None else torch.cat([prev, x], dim=2)
            lead = pad_t - (0 if prev is None else prev.shape[2])
            self.cache[name] = cache_x
        else:
            xin, lead = x, pad_t
        xin = F.pad(xin, (kw // 2, kw // 2, kh // 2, kh // 2, lead, 0))
        return _r(F.conv3d(xin, w, b), self.policy)

    def rms(self, name, x):                                     # RMS_norm :50-64 (channel axis 1)
        g = self.sd[name + ".gamma"]
        g = g.view(1, -1, *([1] * (x.dim() - 2)))
        return _r(F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * g, self.policy)

    def res(self, name, x):                                     # ResidualBlock :198-240
        h = self.cconv(name + ".shortcut", x, cached=False) if (name + ".shortcut.weight") in self.sd else x
        y = _r(F.silu(self.rms(name + ".residual.0", x)), self.policy)
        y = self.cconv(name + ".residual.2", y)
        y = _r(F.silu(self.rms(name + ".residual.3", y)), self.policy)
        y = self.cconv(name + ".residual.6", y)
        return _r(y + h, self.policy)

    def attn(self, name, x):                                    # AttentionBlock :243-282 (one head of width C, per frame)
        b, c, t, h, w = x.shape
        u = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        y = self.rms(name + ".norm", u)
        qkv = _r(F.conv2d(y, self.sd[name + ".to_qkv.weight"], self.sd[name + ".to_qkv.bias"]), self.policy)
        q, k, v = qkv.reshape(b * t, 1, 3 * c, h * w).permute(0, 1, 3, 2).chunk(3, dim=-1)
        o = _r(F.scaled_dot_product_attention(q, k, v), self.policy)
        o = o.squeeze(1).permute(0, 2, 1).reshape(b * t, c, h, w)
        o = _r(F.conv2d(o, self.sd[name + ".proj.weight"], self.sd[name + ".proj.bias"]), self.policy)
        o = o.reshape(b, t, c, h, w).permute(0, 2, 1, 3, 4)
        return _r(o + x, self.policy)

    def resample(self, name, x, temporal):                      # Resample :117-160
        b, c, t, h, w = x.shape
        if temporal:
            key = name + ".time_conv"
            prev = self.cache.get(key)
            if prev is None:
                self.cache[key] = "Rep"                          # first chunk: no temporal up-sampling
            else:
                cache_x = x[:, :, -CACHE_T:].clone()
                if cache_x.shape[2] < 2:
                    head = torch.zeros_like(cache_x) if isinstance(prev, str) else prev[:, :, -1:]
                    cache_x = torch.cat([head, cache_x], dim=2)
                wt, bt = self.sd[key + ".weight"], self.sd[key + ".bias"]
                xin = x if isinstance(prev, str) else torch.cat([prev, x], dim=2)
                xin = F.pad(xin, (0, 0, 0, 0, 2 - (0 if isinstance(prev, str) else prev.shape[2]), 0))
                y = _r(F.conv3d(xin, wt, bt), self.policy)
                self.cache[key] = cache_x
                y = y.reshape(b, 2, c, t, h, w)
                x = torch.stack((y[:, 0], y[:, 1]), 3).reshape(b, c, t * 2, h, w)
                t = t * 2
        u = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        u = F.interpolate(u.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(u)          # Upsample :67-73
        u = _r(F.conv2d(u, self.sd[name + ".resample.1.weight"], self.sd[name + ".resample.1.bias"], padding=1), self.policy)
        return u.reshape(b, t, c, 2 * h, 2 * w).permute(0, 2, 1, 3, 4)


def _run_downsample(self, name, x, temporal):                   # Resample downsample2d / downsample3d :117-160
    b, c, t, h, w = x.shape
    u = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
    u = F.pad(u, (0, 1, 0, 1))                                  # ZeroPad2d((0, 1, 0, 1)) :101-107
    u = _r(F.conv2d(u, self.sd[name + ".resample.1.weight"], self.sd[name + ".resample.1.bias"], stride=2), self.policy)
    x = u.reshape(b, t, c, h // 2, w // 2).permute(0, 2, 1, 3, 4)
    if temporal:
        key = name + ".time_conv"
        prev = self.cache.get(key)
        if prev is None:
            self.cache[key] = x.clone()                        # first chunk: kept as history, no temporal stride :147-149
        else:
            cache_x = x[:, :, -1:].clone()
            xin = torch.cat([prev[:, :, -1:], x], dim=2)
            x = _r(F.conv3d(xin, self.sd[key + ".weight"], self.sd[key + ".bias"], stride=(2, 1, 1)), self.policy)
            self.cache[key] = cache_x
    return x


_Run.downsample = _run_downsample


def dup_up3d(x, out_ch, factor_t, factor_s, first_chunk):      # DupUp3D :395-417
    factor = factor_t * factor_s * factor_s
    rep = out_ch * factor // x.shape[1]
    x = x.repeat_interleave(rep, dim=1)
    x = x.view(x.size(0), out_ch, factor_t, factor_s, factor_s, x.size(2), x.size(3), x.size(4))
    x = x.permute(0, 1, 5, 2, 6, 3, 7, 4).contiguous()
    x = x.view(x.size(0), out_ch, x.size(2) * factor_t, x.size(4) * factor_s, x.size(6) * factor_s)
    return x[:, :, factor_t - 1:] if first_chunk else x


def avg_down3d(x, out_ch, factor_t, factor_s):                  # AvgDown3D :340-372
    pad_t = (factor_t - x.shape[2] % factor_t) % factor_t
    x = F.pad(x, (0, 0, 0, 0, pad_t, 0))
    B, C, T, H, W = x.shape
    factor = factor_t * factor_s * factor_s
    x = x.view(B, C, T // factor_t, factor_t, H // factor_s, factor_s, W // factor_s, factor_s)
    x = x.permute(0, 1, 3, 5, 7, 2, 4, 6).contiguous()
    x = x.view(B, C * factor, T // factor_t, H // factor_s, W // factor_s)
    x = x.view(B, out_ch, C * factor // out_ch, T // factor_t, H // factor_s, W // factor_s)
    return x.mean(dim=2)


def encoder_dims(cfg: dict) -> List[int]:
    return [cfg["enc_dim"] * u for u in [1] + list(cfg["dim_mult"])]      # :527


def encoder_param_specs(cfg: dict):
    """[(key, shape, std, mean)] of the encoder half of AutoencoderKLWan2_2_ (encoder + conv1, :505-562, :771)."""
    z, dims = cfg["z_dim"], encoder_dims(cfg)
    t_dn = list(cfg["temperal_downsample"])
    specs = []

    def conv(name, co, ci, k):
        fan = ci * math.prod(k)
        specs.append((name + ".weight", (co, ci) + tuple(k), fan ** -0.5, 0.0))
        specs.append((name + ".bias", (co,), 0.02, 0.0))

    def res(name, ci, co):
        specs.append((name + ".residual.0.gamma", (ci, 1, 1, 1), 0.1, 1.0))
        conv(name + ".residual.2", co, ci, (3, 3, 3))
        specs.append((name + ".residual.3.gamma", (co, 1, 1, 1), 0.1, 1.0))
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
    specs.append(("encoder.middle.1.norm.gamma", (dims[-1], 1, 1), 0.1, 1.0))
    conv("encoder.middle.1.to_qkv", 3 * dims[-1], dims[-1], (1, 1))
    conv("encoder.middle.1.proj", dims[-1], dims[-1], (1, 1))
    res("encoder.middle.2", dims[-1], dims[-1])
    specs.append(("encoder.head.0.gamma", (dims[-1], 1, 1, 1), 0.1, 1.0))
    conv("encoder.head.2", 2 * z, dims[-1], (3, 3, 3))
    conv("conv1", 2 * z, 2 * z, (1, 1, 1))
    return specs


def encoder_state_dict(cfg: dict, tag: str = "vae"):
    from oracle import synth
    return {name: synth.tensor(f"{tag}/{name}", shape, std, mean) for name, shape, std, mean in encoder_param_specs(cfg)}


def encoder_state_dict_torch(cfg: dict, device, dtype=None, tag: str = "vae"):
    from oracle import synth
    return {name: synth.tensor_torch(f"{tag}/{name}", shape, std, mean, device=device, dtype=dtype)
            for name, shape, std, mean in encoder_param_specs(cfg)}


def video(cfg: dict, T: int, H: int, W: int, tag: str = "vaex"):
    """Synthetic pixel video [1, 3, T, H, W] in [-1, 1]."""
    from oracle import synth
    return np_clip(synth.tensor(f"{tag}/video", (1, 3, T, H, W), 0.5))


def np_clip(a):
    import numpy as np
    return np.clip(a, -1.0, 1.0)


def encode(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, scale, policy: str = "fp32") -> torch.Tensor:
    """AutoencoderKLWan2_2_.encode (:788-819): x [1, 3, T, H, W] (T = 1 + 4k) -> [1, 2*z_dim, 1 + k, H/16, W/16] =
    normalised mean | log-variance. Patchify, first frame alone then 4 frames per chunk through the encoder with its
    feature cache, conv1, (mu - mean) * (1 / std)."""
    run = _Run(sd, policy)
    dims = encoder_dims(cfg)
    t_dn = list(cfg["temperal_downsample"])
    n = len(cfg["dim_mult"])
    zd = cfg["z_dim"]
    b, c, f, h, w = x.shape                                                                # patchify :285-301
    x = x.view(b, c, f, h // 2, 2, w // 2, 2).permute(0, 1, 6, 4, 2, 3, 5).reshape(b, c * 4, f, h // 2, w // 2)
    outs = []
    for i in range(1 + (f - 1) // 4):                                                       # :795-810
        xc = x[:, :, :1] if i == 0 else x[:, :, 1 + 4 * (i - 1):1 + 4 * i]
        y = run.cconv("encoder.conv1", xc)
        for s in range(n):                                                                  # Down_ResidualBlock :452-457
            name = f"encoder.downsamples.{s}.downsamples."
            down = s != n - 1
            temporal = down and s < len(t_dn) and bool(t_dn[s])
            y_in = y
            for j in range(cfg["num_res_blocks"]):
                y = run.res(name + str(j), y)
            if down:
                y = run.downsample(name + str(cfg["num_res_blocks"]), y, temporal)
            y = _r(y + avg_down3d(y_in, dims[s + 1], 2 if temporal else 1, 2 if down else 1), policy)
        y = run.res("encoder.middle.0", y)
        y = run.attn("encoder.middle.1", y)
        y = run.res("encoder.middle.2", y)
        y = _r(F.silu(run.rms("encoder.head.0", y)), policy)
        outs.append(run.cconv("encoder.head.2", y))
    out = run.cconv("conv1", torch.cat(outs, dim=2), cached=False)                          # :811
    mu, log_var = out.chunk(2, dim=1)
    mu = _r((mu - scale[0].view(1, zd, 1, 1, 1).to(mu)) * scale[1].view(1, zd, 1, 1, 1).to(mu), policy)   # :812-816
    return torch.cat([mu, log_var], dim=1)


def decode(sd: Dict[str, torch.Tensor], cfg: dict, z: torch.Tensor, scale, policy: str = "fp32") -> torch.Tensor:
    """z: [1, z_dim, T, H, W] normalised latents -> video [1, 3, 1 + 4 (T-1), 16 H, 16 W] in [-1, 1] (fp32)."""
    run = _Run(sd, policy)
    dims = decoder_dims(cfg)
    t_up = cfg["temperal_downsample"][::-1]
    n = len(cfg["dim_mult"])
    zd = cfg["z_dim"]
    z = _r(z / scale[1].view(1, zd, 1, 1, 1).to(z) + scale[0].view(1, zd, 1, 1, 1).to(z), policy)       # :824-829
    x_all = run.cconv("conv2", z, cached=False)                                                          # :831
    outs = []
    for i in range(z.shape[2]):                                                                          # :832-846
        first = i == 0
        x = run.cconv("decoder.conv1", x_all[:, :, i:i + 1])
        x = run.res("decoder.middle.0", x)
        x = run.attn("decoder.middle.1", x)
        x = run.res("decoder.middle.2", x)
        for s in range(n):                                                                               # Up_ResidualBlock
            name = f"decoder.upsamples.{s}.upsamples."
            up = s != n - 1
            temporal = up and s < len(t_up) and bool(t_up[s])
            main = x
            for j in range(cfg["num_res_blocks"] + 1):
                main = run.res(name + str(j), main)
            if up:
                main = run.resample(name + str(cfg["num_res_blocks"] + 1), main, temporal)
                short = dup_up3d(x, dims[s + 1], 2 if temporal else 1, 2, first)
                x = _r(main + short, policy)
            else:
                x = main
        y = _r(F.silu(run.rms("decoder.head.0", x)), policy)
        outs.append(run.cconv("decoder.head.2", y))
    out = torch.cat(outs, dim=2)
    b, c, f, h, w = out.shape                                                                            # unpatchify :304-318
    out = out.view(b, c // 4, 2, 2, f, h, w).permute(0, 1, 4, 5, 3, 6, 2).reshape(b, c // 4, f, h * 2, w * 2)
    return out.clamp(-1, 1)                                                                              # :1043
