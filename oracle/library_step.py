"""TEST / MEASUREMENT INFRASTRUCTURE — the reference's LIBRARY path for the denoising step, restated in stock torch.

What the unmodified reference executes on a GPU is torch modules under ``torch.autocast('cuda', bf16)``: cuBLAS
``nn.Linear``, cuDNN ``Conv3d``, ``F.layer_norm``/``group_norm`` in fp32, complex128 RoPE with a Python loop over the
batch, the fp32 per-token time MLP, and ``flash_attn_varlen_func`` (flash-attn 2) or
``F.scaled_dot_product_attention`` behind ``attention()`` (FlexAM/models/attention_utils.py:174-233). The reference
sources cannot travel to the GPU box, so this file restates that module tree (same parameter names as the reference
state_dict, same op order and dtype flow; every stage cites FlexAM/models/wan_transformer3d_FlexAM.py as :line) with
nothing but library calls. It serves two purposes, both on the checker side:

  * ``bench.py``'s ``library_baseline`` leg: the step time of the reference's own GPU path on the same B200
    (SURVEY.md §8d "GPU reference number"), CUDA-event timed, with per-family (Linear / attention) times;
  * a second parity anchor: the native output against the real autocast numerics of cuBLAS + flash-attn.

Never imported by ``flexam_b200``. Checked against the live reference module on CPU (fp32) in tests/test_oracle.py.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


def sinusoid(dim: int, pos: torch.Tensor) -> torch.Tensor:          # :31-41
    half = dim // 2
    s = torch.outer(pos.to(torch.float64), torch.pow(10000, -torch.arange(half, device=pos.device).to(torch.float64).div(half)))
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1)


def rope_freqs(head_dim: int, max_len: int = 1024, theta: float = 10000.0) -> torch.Tensor:   # :44-52, :655-665
    d = head_dim
    parts = []
    for dim in (d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)):
        ang = torch.outer(torch.arange(max_len), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float64).div(dim)))
        parts.append(torch.polar(torch.ones_like(ang), ang))
    return torch.cat(parts, dim=1)


@torch.autocast("cuda", enabled=False)
def rope_apply(x: torch.Tensor, grid, freqs: torch.Tensor) -> torch.Tensor:
    """:135-164 — per-sample Python loop, complex64 x complex128 product, cast back to x.dtype."""
    n, c = x.size(2), x.size(3) // 2
    fr = freqs.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    f, h, w = grid
    seq = f * h * w
    out = []
    for i in range(x.size(0)):
        xi = torch.view_as_complex(x[i, :seq].to(torch.float32).reshape(seq, n, -1, 2))
        fi = torch.cat([fr[0][:f].view(f, 1, 1, -1).expand(f, h, w, -1), fr[1][:h].view(1, h, 1, -1).expand(f, h, w, -1),
                        fr[2][:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, -1)
        xi = torch.view_as_real(xi * fi).flatten(2)
        out.append(torch.cat([xi, x[i, seq:]]))
    return torch.stack(out).to(x.dtype)


class RMSNorm(nn.Module):                                            # :173-189
    def __init__(self, dim, eps):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return x * torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + self.eps).to(x.dtype) * self.weight


class Attention(nn.Module):                                          # :205-262 (self), :353-371 (cross)
    def __init__(self, dim, heads, eps, backend):
        super().__init__()
        self.heads, self.hd, self.backend = heads, dim // heads, backend
        self.q, self.k, self.v, self.o = (nn.Linear(dim, dim) for _ in range(4))
        self.norm_q, self.norm_k = RMSNorm(dim, eps), RMSNorm(dim, eps)

    def attend(self, q, k, v, dtype):
        """attention() :174-233: flash-attn 2 varlen (the default when flash_attn imports) or SDPA."""
        q, k, v = q.to(dtype), k.to(dtype), v.to(dtype)
        if self.backend == "flash":
            from flash_attn import flash_attn_varlen_func
            b, lq, lk = q.size(0), q.size(1), k.size(1)
            cq = torch.arange(0, (b + 1) * lq, lq, dtype=torch.int32, device=q.device)
            ck = torch.arange(0, (b + 1) * lk, lk, dtype=torch.int32, device=q.device)
            o = flash_attn_varlen_func(q.flatten(0, 1), k.flatten(0, 1), v.flatten(0, 1), cq, ck, lq, lk)
            return o.unflatten(0, (b, lq))
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
        return o.transpose(1, 2).contiguous()

    def forward(self, x, dtype, context=None, grid=None, freqs=None):
        b, n, d = x.size(0), self.heads, self.hd
        src = x.to(dtype) if context is None else context.to(dtype)
        q = self.norm_q(self.q(x.to(dtype))).view(b, -1, n, d)
        k = self.norm_k(self.k(src)).view(b, -1, n, d)
        v = self.v(src).view(b, -1, n, d)
        if context is None:
            q, k = rope_apply(q, grid, freqs), rope_apply(k, grid, freqs)
        return self.o(self.attend(q, k, v, dtype).to(dtype).flatten(2))


class Block(nn.Module):                                              # :376-472
    def __init__(self, dim, ffn_dim, heads, eps, backend):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.self_attn = Attention(dim, heads, eps, backend)
        self.norm3 = nn.LayerNorm(dim, eps, elementwise_affine=True)
        self.cross_attn = Attention(dim, heads, eps, backend)
        self.norm2 = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.ffn = nn.Sequential(nn.Linear(dim, ffn_dim), nn.GELU(approximate="tanh"), nn.Linear(ffn_dim, dim))
        self.modulation = nn.Parameter(torch.zeros(1, 6, dim))
        self.modulation_density = nn.Parameter(torch.zeros(1, 2, dim))

    def forward(self, x, e, dens, grid, freqs, context, dtype):
        if e.dim() > 3:                                               # per-token modulation :444-446
            e = [u.squeeze(2) for u in (self.modulation.unsqueeze(0) + e).chunk(6, dim=2)]
        else:
            e = (self.modulation + e).chunk(6, dim=1)
        dens = (self.modulation_density + dens).chunk(2, dim=1)
        h = (self.norm1(x) * (1 + e[1]) + e[0] + dens[0]).to(dtype)
        x = x + self.self_attn(h, dtype, grid=grid, freqs=freqs) * e[2]
        x = x + self.cross_attn(self.norm3(x), dtype, context=context)
        h = (self.norm2(x) * (1 + e[4]) + e[3] + dens[1]).to(dtype)
        return x + self.ffn(h) * e[5]


class Head(nn.Module):                                               # :475-507
    def __init__(self, dim, out, eps):
        super().__init__()
        self.norm = nn.LayerNorm(dim, eps, elementwise_affine=False)
        self.head = nn.Linear(dim, out)
        self.modulation = nn.Parameter(torch.zeros(1, 2, dim))
        self.modulation_density = nn.Parameter(torch.zeros(1, 1, dim))

    def forward(self, x, e, dens):
        if e.dim() > 2:
            e = [u.squeeze(2) for u in (self.modulation.unsqueeze(0) + e.unsqueeze(2)).chunk(2, dim=2)]
        else:
            e = (self.modulation + e.unsqueeze(1)).chunk(2, dim=1)
        d = self.modulation_density + dens.unsqueeze(1)
        return self.head(self.norm(x) * (1 + e[1]) + e[0] + d)


def _cnn_stage(ci, co, groups):
    return nn.Sequential(nn.Conv3d(ci, co, (1, 3, 3), padding=(0, 1, 1)), nn.GroupNorm(groups, co), nn.SiLU())


class LibraryStep(nn.Module):
    """The FlexAM transformer as the reference builds it (:526-728, :1335-1438), library kernels only."""

    def __init__(self, cfg: dict, backend: str = "flash"):
        super().__init__()
        D, eps = cfg["dim"], cfg["eps"]
        self.cfg, self.backend = dict(cfg), backend
        self.patch_embedding = nn.Conv3d(cfg["in_dim"], D, tuple(cfg["patch_size"]), stride=tuple(cfg["patch_size"]))
        self.text_embedding = nn.Sequential(nn.Linear(cfg["text_dim"], D), nn.GELU(approximate="tanh"), nn.Linear(D, D))
        self.time_embedding = nn.Sequential(nn.Linear(cfg["freq_dim"], D), nn.SiLU(), nn.Linear(D, D))
        self.time_projection = nn.Sequential(nn.SiLU(), nn.Linear(D, 6 * D))
        self.density_embedding = nn.Sequential(nn.Linear(cfg["freq_dim"], D), nn.SiLU(), nn.Linear(D, D))
        self.density_projection = nn.Sequential(nn.SiLU(), nn.Linear(D, 2 * D))
        self.blocks = nn.ModuleList(Block(D, cfg["ffn_dim"], cfg["num_heads"], eps, backend)
                                    for _ in range(cfg["num_layers"]))
        self.head = Head(D, cfg["out_dim"] * math.prod(cfg["patch_size"]), eps)
        self.ref_conv = nn.Conv2d(cfg.get("in_dim_ref_conv", cfg["out_dim"]), D, tuple(cfg["patch_size"][1:]),
                                  stride=tuple(cfg["patch_size"][1:]))
        ci = cfg.get("in_dim_cnn_block", cfg.get("in_dim_cnn"))
        self.cnn_conv1, self.cnn_conv2 = _cnn_stage(ci, 192, 24), _cnn_stage(192, 192, 24)
        self.cnn_conv3, self.cnn_conv4 = _cnn_stage(192, 96, 12), _cnn_stage(96, 96, 12)
        self.cnn_conv5 = nn.Conv3d(96, cfg.get("out_dim_cnn_block", cfg.get("out_dim_cnn")), 1)
        self.freqs = rope_freqs(D // cfg["num_heads"])
        for p in self.parameters():
            p.requires_grad_(False)

    @torch.no_grad()
    def forward(self, x, t, context: List[torch.Tensor], seq_len: int, y, full_ref, additional_control, density):
        """forward() :817-1123 for the FlexAM inputs, sp_world_size 1, no TeaCache. Call under
        ``torch.autocast('cuda', dtype=torch.bfloat16)`` like the pipeline does (pipeline :503, :912-923)."""
        cfg = self.cfg
        dtype = x.dtype
        dev = x.device
        if self.freqs.device != dev:
            self.freqs = self.freqs.to(dev)
        C = x.shape[1]
        u = torch.cat([y[:, :C], additional_control], dim=1)                                   # :868-881
        x1 = self.cnn_conv1(u)
        x2 = self.cnn_conv2(x1) + x1
        x3 = self.cnn_conv3(x2)
        x4 = self.cnn_conv4(x3) + x3
        yy = torch.cat([self.cnn_conv5(x4), y[:, C:]], dim=1)
        xs = [self.patch_embedding(torch.cat([a, b], dim=0).unsqueeze(0)) for a, b in zip(x, yy)]   # :883-885
        F_, Hp, Wp = xs[0].shape[2:]
        xs = [a.flatten(2).transpose(1, 2) for a in xs]
        ref = self.ref_conv(full_ref).flatten(2).transpose(1, 2)                               # :895-899
        grid = (F_ + 1, Hp, Wp)
        seq_len = seq_len + ref.size(1)
        xs = torch.cat([torch.cat([r.unsqueeze(0), a], dim=1) for r, a in zip(ref, xs)])
        if t.dim() != 1 and t.size(1) < seq_len:                                               # :900-904
            t = torch.cat([t[:, -1:].repeat(1, seq_len - t.size(1)), t], dim=1)
        with torch.autocast("cuda", dtype=torch.float32):                                      # :928-955
            if t.dim() != 1:
                bt = t.size(0)
                e = self.time_embedding(sinusoid(cfg["freq_dim"], t.flatten()).unflatten(0, (bt, seq_len)).float())
                e0 = self.time_projection(e).unflatten(2, (6, cfg["dim"]))
            else:
                e = self.time_embedding(sinusoid(cfg["freq_dim"], t).float())
                e0 = self.time_projection(e).unflatten(1, (6, cfg["dim"]))
            de = self.density_embedding(sinusoid(cfg["freq_dim"], density).float())
            de0 = self.density_projection(de).unflatten(1, (2, cfg["dim"]))
        ctx = self.text_embedding(torch.stack(                                                 # :958-964
            [torch.cat([c, c.new_zeros(cfg["text_len"] - c.size(0), c.size(1))]) for c in context]))
        for blk in self.blocks:
            xs = blk(xs, e0, de0, grid, self.freqs, ctx, dtype)
        o = self.head(xs, e, de)[:, ref.size(1):]                                               # :1101-1109
        Cc = cfg["out_dim"]
        o = o.view(o.size(0), F_, Hp, Wp, 1, 2, 2, Cc)
        return torch.einsum("bfhwpqrc->bcfphqwr", o).reshape(o.size(0), Cc, F_, Hp * 2, Wp * 2)   # :1126-1149


def flash_attn_usable(device) -> Optional[str]:
    """None when flash-attn 2 runs on this device, else the reason (import error / no kernel image for the arch)."""
    try:
        from flash_attn import flash_attn_func
        q = torch.zeros(1, 128, 1, 128, device=device, dtype=torch.bfloat16)
        flash_attn_func(q, q, q)
        torch.cuda.synchronize(device)
        return None
    except Exception as exc:   # noqa: BLE001 — any failure means "not usable here"
        return f"{type(exc).__name__}: {str(exc)[:120]}"
