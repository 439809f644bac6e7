"""TEST INFRASTRUCTURE — deterministic synthetic weights and inputs for the FlexAM DiT (no RNG library state).

Every tensor is a pure function of (name, shape): a counter-based 64-bit integer hash (splitmix64 finaliser) mapped
to a centred uniform and scaled, then rounded to a bf16-representable fp32 value so the fp32 oracle, the reference
module and the bf16 native path all see bit-identical parameters. Integer arithmetic only => identical on every
numpy version / machine, which is what lets ``tests/golden/*.npz`` hold outputs only.

Parameter names and shapes are the reference's state_dict (wan_transformer3d_FlexAM.py:624-711,
:402-420, :475-491).
"""
from __future__ import annotations

import zlib

import numpy as np

CONFIGS = {
    # Wan2.2-Fun-5B FlexAM (dims from the HF checkpoint config)
    "real": dict(dim=3072, ffn_dim=14336, num_heads=24, num_layers=30, in_dim=148, out_dim=48, text_len=512,
                 text_dim=4096, freq_dim=256, eps=1e-6, patch_size=(1, 2, 2), in_dim_cnn=288, out_dim_cnn=48),
    # real width, 2 layers: exercises every kernel at production tile shapes, CPU-runnable in seconds
    "real2": dict(dim=3072, ffn_dim=14336, num_heads=24, num_layers=2, in_dim=148, out_dim=48, text_len=512,
                  text_dim=4096, freq_dim=256, eps=1e-6, patch_size=(1, 2, 2), in_dim_cnn=288, out_dim_cnn=48),
    # 2 heads x 128: smallest shape the native kernels accept
    "tiny": dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, in_dim=148, out_dim=48, text_len=512,
                 text_dim=64, freq_dim=256, eps=1e-6, patch_size=(1, 2, 2), in_dim_cnn=288, out_dim_cnn=48),
}


def _splitmix64(x: np.ndarray) -> np.ndarray:
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        x = x ^ (x >> np.uint64(31))
    return x


def to_bf16_f32(a: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest-even bf16 value, returned as fp32."""
    u = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + np.uint64(0x7FFF) + ((u >> np.uint64(16)) & np.uint64(1))) & np.uint64(0xFFFF0000)
    return u.astype(np.uint32).view(np.float32).reshape(a.shape)


def tensor(name: str, shape, std: float = 1.0, mean: float = 0.0, bf16: bool = True, chunk: int = 1 << 24) -> np.ndarray:
    """Centred uniform with the requested std (|x - mean| <= std*sqrt(3)), element i = hash(seed(name), i)."""
    n = int(np.prod(shape))
    seed = np.uint64((zlib.crc32(name.encode("utf-8")) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF)
    out = np.empty(n, dtype=np.float32)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        with np.errstate(over="ignore"):
            h = _splitmix64(np.arange(s, e, dtype=np.uint64) + seed)
        u = (h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))  # [0,1)
        out[s:e] = (u - np.float32(0.5)) * np.float32(2.0 * np.sqrt(3.0) * std) + np.float32(mean)
    out = out.reshape(shape)
    return to_bf16_f32(out) if bf16 else out


def param_specs(cfg: dict):
    """[(state_dict key, shape, std, mean)] for the reference module built with ``cfg``."""
    D, Fd, L = cfg["dim"], cfg["ffn_dim"], cfg["num_layers"]
    specs = []

    def lin(prefix, out_f, in_f, wstd=None):
        specs.append((prefix + ".weight", (out_f, in_f), wstd if wstd is not None else in_f ** -0.5, 0.0))
        specs.append((prefix + ".bias", (out_f,), 0.02, 0.0))

    pt, ph, pw = cfg["patch_size"]
    specs.append(("patch_embedding.weight", (D, cfg["in_dim"], pt, ph, pw), (cfg["in_dim"] * pt * ph * pw) ** -0.5, 0.0))
    specs.append(("patch_embedding.bias", (D,), 0.02, 0.0))
    lin("text_embedding.0", D, cfg["text_dim"]); lin("text_embedding.2", D, D)
    lin("time_embedding.0", D, cfg["freq_dim"]); lin("time_embedding.2", D, D)
    lin("time_projection.1", 6 * D, D)
    lin("density_embedding.0", D, cfg["freq_dim"]); lin("density_embedding.2", D, D)
    lin("density_projection.1", 2 * D, D)
    for i in range(L):
        b = f"blocks.{i}."
        specs.append((b + "modulation", (1, 6, D), 0.3, 0.0))
        specs.append((b + "modulation_density", (1, 2, D), 0.3, 0.0))
        for att in ("self_attn", "cross_attn"):
            for nm in ("q", "k", "v", "o"):
                lin(b + f"{att}.{nm}", D, D)
            specs.append((b + f"{att}.norm_q.weight", (D,), 0.1, 1.0))
            specs.append((b + f"{att}.norm_k.weight", (D,), 0.1, 1.0))
        specs.append((b + "norm3.weight", (D,), 0.1, 1.0))
        specs.append((b + "norm3.bias", (D,), 0.05, 0.0))
        lin(b + "ffn.0", Fd, D); lin(b + "ffn.2", D, Fd)
    lin("head.head", cfg["out_dim"] * pt * ph * pw, D)
    specs.append(("head.modulation", (1, 2, D), 0.3, 0.0))
    specs.append(("head.modulation_density", (1, 1, D), 0.3, 0.0))
    specs.append(("ref_conv.weight", (D, cfg["out_dim"], ph, pw), (cfg["out_dim"] * ph * pw) ** -0.5, 0.0))
    specs.append(("ref_conv.bias", (D,), 0.02, 0.0))
    chans = [(cfg["in_dim_cnn"], 192), (192, 192), (192, 96), (96, 96)]
    for j, (ci, co) in enumerate(chans, start=1):
        specs.append((f"cnn_conv{j}.0.weight", (co, ci, 1, 3, 3), (ci * 9) ** -0.5, 0.0))
        specs.append((f"cnn_conv{j}.0.bias", (co,), 0.02, 0.0))
        specs.append((f"cnn_conv{j}.1.weight", (co,), 0.1, 1.0))
        specs.append((f"cnn_conv{j}.1.bias", (co,), 0.05, 0.0))
    specs.append(("cnn_conv5.weight", (cfg["out_dim_cnn"], 96, 1, 1, 1), 96 ** -0.5, 0.0))
    specs.append(("cnn_conv5.bias", (cfg["out_dim_cnn"],), 0.02, 0.0))
    return specs


def state_dict(cfg: dict, tag: str = "w"):
    """name -> np.float32 array (bf16-representable values)."""
    return {name: tensor(f"{tag}/{name}", shape, std, mean) for name, shape, std, mean in param_specs(cfg)}


def tensor_torch(name: str, shape, std: float = 1.0, mean: float = 0.0, device="cpu", dtype=None,
                 chunk: int = 1 << 24):
    """``tensor()`` evaluated with torch on ``device``: bit-identical values (tests/test_oracle.py), so that the 5 B
    parameters of the full-depth model can be generated on the GPU in seconds instead of minutes of numpy hashing.
    uint64 arithmetic is done in int64 (two's-complement wrap-around is the same ring; logical right shifts are
    arithmetic shifts with the sign extension masked off). Returns fp32 (bf16-representable) or ``dtype``."""
    import torch

    def i64(c: int) -> int:          # a 64-bit constant as the int64 with the same bit pattern
        c &= 0xFFFFFFFFFFFFFFFF
        return c - (1 << 64) if c >= (1 << 63) else c

    def lsr(x, k: int):
        return (x >> k) & ((1 << (64 - k)) - 1)

    n = int(np.prod(shape))
    seed = i64(zlib.crc32(name.encode("utf-8")) * 0x100000001B3)
    scale = float(np.float32(2.0 * np.sqrt(3.0) * std))
    out = torch.empty(n, dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        x = torch.arange(s, e, dtype=torch.int64, device=device) + seed
        x = x + i64(0x9E3779B97F4A7C15)
        x = (x ^ lsr(x, 30)) * i64(0xBF58476D1CE4E5B9)
        x = (x ^ lsr(x, 27)) * i64(0x94D049BB133111EB)
        x = x ^ lsr(x, 31)
        u = lsr(x, 40).to(torch.float32) * float(np.float32(1.0 / (1 << 24)))
        out[s:e] = (u - 0.5) * scale + float(np.float32(mean))
    out = out.reshape(tuple(shape)).to(torch.bfloat16)       # round to nearest even, like to_bf16_f32
    return out.to(dtype if dtype is not None else torch.float32)


def state_dict_torch(cfg: dict, device, dtype=None, tag: str = "w"):
    """``state_dict()`` generated on ``device`` (see tensor_torch)."""
    return {name: tensor_torch(f"{tag}/{name}", shape, std, mean, device=device, dtype=dtype)
            for name, shape, std, mean in param_specs(cfg)}


def inputs(cfg: dict, F: int, H: int, W: int, B: int = 2, per_token_t: bool = True, tag: str = "in",
           prompt_lens=(37, 120), t_value: float = 875.0, density: float = 0.1):
    """Synthetic forward() arguments on the latent grid (F, H, W)."""
    C = cfg["out_dim"]
    x = tensor(f"{tag}/x", (B, C, F, H, W))
    y = tensor(f"{tag}/y", (B, cfg["in_dim"] - C, F, H, W))
    mask = np.ones((F, H, W), np.float32)
    mask[0] = 0.0
    y[:, C:C + 4] = mask  # mask channels in {0,1}: first latent frame pinned
    add = tensor(f"{tag}/additional_control", (B, cfg["in_dim_cnn"] - C, F, H, W))
    full_ref = tensor(f"{tag}/full_ref", (B, C, H, W))
    context = [tensor(f"{tag}/context{i}", (prompt_lens[i % len(prompt_lens)], cfg["text_dim"])) for i in range(B)]
    L0 = F * (H // 2) * (W // 2)
    if per_token_t == "frac":
        # fg/bg-edit regime (pipeline :686-690, :891-898): the latent mask is a TRILINEAR resize of the pixel mask kept
        # in bf16, so boundary tokens carry fractional factors and the per-token timesteps `mask * t` take many distinct
        # values. Synthetic stand-in: a smooth ramp in [0, 1] over (f, h, w), rounded to bf16, times t (bf16 product).
        f, h, w = np.meshgrid(np.arange(F), np.arange(H // 2), np.arange(W // 2), indexing="ij")
        ramp = (0.15 * f / max(F - 1, 1) + 0.55 * h / max(H // 2 - 1, 1) + 0.30 * w / max(W // 2 - 1, 1)).astype(np.float32)
        ramp = np.clip(1.25 * ramp - 0.1, 0.0, 1.0)
        t_tok = to_bf16_f32(to_bf16_f32(ramp.reshape(-1)) * np.float32(t_value))
        t = np.broadcast_to(t_tok, (B, L0)).copy()
    elif per_token_t:
        t_tok = (mask[:, ::2, ::2].reshape(-1) * np.float32(t_value)).astype(np.float32)
        t = np.broadcast_to(t_tok, (B, L0)).copy()
    else:
        t = np.full((B,), t_value, np.float32)
    dens = np.full((B,), density, np.float32)
    return dict(x=x, y=y, additional_control=add, full_ref=full_ref, context=context, t=t, density=dens, seq_len=L0)


def loop_inputs(cfg: dict, F: int, H: int, W: int, tag: str = "loop", prompt_lens=(37, 120)):
    """Synthetic arguments of the sampling loop (pipeline_wan2_2_fun_control_FlexAM.py:843-934) for ONE video:
    noisy latents, the latent mask with the first frame pinned, masked-video / mask / control / additional-control
    latents, the reference-image latents and the negative / positive prompt embeddings."""
    C = cfg["out_dim"]
    mask = np.ones((1, 1, F, H, W), np.float32)
    mask[:, :, 0] = 0.0
    return dict(
        latents=tensor(f"{tag}/latents", (1, C, F, H, W)),
        mask=mask,
        masked_video_latents=tensor(f"{tag}/masked_video", (1, C, F, H, W)),
        mask_latents=np.broadcast_to(1.0 - mask, (1, 4, F, H, W)).astype(np.float32).copy(),
        control_video_latents=tensor(f"{tag}/control", (1, C, F, H, W)),
        additional_control=tensor(f"{tag}/additional_control", (1, cfg["in_dim_cnn"] - C, F, H, W)),
        ref_image_latents=tensor(f"{tag}/ref", (1, C, H, W)),
        negative_prompt_embeds=[tensor(f"{tag}/neg", (prompt_lens[0], cfg["text_dim"]))],
        prompt_embeds=[tensor(f"{tag}/pos", (prompt_lens[1], cfg["text_dim"]))],
    )
