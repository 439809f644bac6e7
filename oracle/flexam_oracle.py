"""TEST INFRASTRUCTURE — CPU restatement of the reference FlexAM denoising step.

This is the ORACLE (checker) for the native path. Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product (``flexam_b200``) never does.

It restates ``Wan2_2Transformer3DModel_FlexAM.forward`` (reference file
``FlexAM/models/wan_transformer3d_FlexAM.py``, cited as :line) as one functional pass over a plain
``{state_dict key: tensor}`` mapping, in plain torch ops.

Pinning: the reference ships no tests or golden vectors for this path, so the pin is the reference
module itself, run in the build container: ``oracle/make_golden.py`` imports the real module (``oracle/ref_import.py``)
and stores its fp32 CPU outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks this file against those
outputs (<= 2e-5 relative L2) and, when /root/reference is present, against the live module.

Precision policies
  "fp32": every op in fp32 — what the reference computes on CPU (autocast is a no-op there).
  "bf16": emulates the reference's CUDA bf16-autocast dtype flow (:237-242 of SURVEY.md) by rounding to bf16 exactly
          where the reference materialises a bf16 tensor: Linear/conv outputs, RMSNorm's three roundings (:186-189),
          RoPE output (:164), attention output, modulated activations cast `.to(dtype)` (:453,:465). The residual
          stream, LayerNorm, modulation and embedding MLPs stay fp32 as under autocast.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F


def _r(t: torch.Tensor, policy: str) -> torch.Tensor:
    """Materialise a tensor in the policy's storage dtype (bf16 round trip) and return it as fp32."""
    return t.to(torch.bfloat16).to(torch.float32) if policy == "bf16" else t


def sinusoidal_embedding_1d(dim: int, position: torch.Tensor) -> torch.Tensor:
    # :31-41 — float64, cos block first
    half = dim // 2
    pos = position.to(torch.float64)
    freqs = torch.pow(10000.0, -torch.arange(half, dtype=torch.float64, device=pos.device) / half)
    s = torch.outer(pos, freqs)
    return torch.cat([torch.cos(s), torch.sin(s)], dim=1)


def rope_angles(head_dim: int = 128, max_len: int = 1024, theta: float = 10000.0) -> torch.Tensor:
    """float64 angle table [max_len, head_dim/2] for the three concatenated axis tables (:44-52, :655-665)."""
    d = head_dim
    parts = []
    for dim in (d - 4 * (d // 6), 2 * (d // 6), 2 * (d // 6)):
        inv = 1.0 / torch.pow(theta, torch.arange(0, dim, 2, dtype=torch.float64) / dim)
        parts.append(torch.outer(torch.arange(max_len, dtype=torch.float64), inv))
    return torch.cat(parts, dim=1)


def rope_table_f32(head_dim: int = 128, max_len: int = 1024) -> torch.Tensor:
    """[max_len, head_dim/2, 2] (cos, sin) fp32 — the table the native kernel consumes."""
    a = rope_angles(head_dim, max_len)
    return torch.stack([a.cos(), a.sin()], dim=-1).to(torch.float32).contiguous()


def rope_apply(x: torch.Tensor, grid: Sequence[int], angles: torch.Tensor) -> torch.Tensor:
    """x: [L, n_heads, head_dim]; rotates the first f*h*w tokens (:135-164)."""
    f, h, w = grid
    n, d = x.shape[1], x.shape[2]
    c = d // 2
    seq = f * h * w
    a0, a1, a2 = angles.split([c - 2 * (c // 3), c // 3, c // 3], dim=1)
    ang = torch.cat([a0[:f].view(f, 1, 1, -1).expand(f, h, w, -1), a1[:h].view(1, h, 1, -1).expand(f, h, w, -1),
                     a2[:w].view(1, 1, w, -1).expand(f, h, w, -1)], dim=-1).reshape(seq, 1, c).to(x.device)
    xr = x[:seq].to(torch.float64).reshape(seq, n, c, 2)
    cos, sin = ang.cos(), ang.sin()
    re = xr[..., 0] * cos - xr[..., 1] * sin
    im = xr[..., 0] * sin + xr[..., 1] * cos
    out = torch.stack([re, im], dim=-1).reshape(seq, n, d).to(torch.float32)
    return torch.cat([out, x[seq:]], dim=0)


def _linear(x, w, b, policy):
    # nn.Linear under autocast: bf16 operands, fp32 accumulate, one rounding of (acc + bias)
    return _r(F.linear(_r(x, policy), w, b), policy)


def _rmsnorm(x, w, eps, policy):
    # :186-189 — mean over the FULL row; under autocast pow/mean/rsqrt are fp32, r is cast to x.dtype
    r = torch.rsqrt(x.pow(2).mean(dim=-1, keepdim=True) + eps)
    return _r(_r(x * _r(r, policy), policy) * w, policy)


def _attention(q, k, v, policy):
    # attention() :174-233 — non-causal softmax(q k^T / sqrt(d)) v ; q,k,v: [L, n, d]
    o = F.scaled_dot_product_attention(q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1))
    return _r(o.transpose(0, 1), policy)


def _gelu_tanh(x):
    return F.gelu(x, approximate="tanh")


def cnn_fuser(sd, y_ctrl0, additional_control, policy):
    """:868-880 with defs :680-711. Inputs [B, C, F, H, W]; returns [B, 48, F, H, W]."""
    def conv(u, name, pad):
        return _r(F.conv3d(_r(u, policy), sd[name + ".weight"], sd[name + ".bias"], padding=pad), policy)

    def stage(u, j, groups):
        c = conv(u, f"cnn_conv{j}.0", (0, 1, 1))
        g = F.group_norm(c, groups, sd[f"cnn_conv{j}.1.weight"], sd[f"cnn_conv{j}.1.bias"], eps=1e-5)
        return F.silu(g)

    u = torch.cat([y_ctrl0, additional_control], dim=1)
    x1 = stage(u, 1, 24)
    x2 = stage(x1, 2, 24) + x1
    x3 = stage(x2, 3, 12)
    x4 = stage(x3, 4, 12) + x3
    return conv(x4, "cnn_conv5", 0)


def embed_mlps(sd, cfg, t_flat, density, policy):
    """fp32 time / density MLPs (:928-955). t_flat: [n]; returns e [n,D], e0 [n,6,D], dens [B,D], dens0 [B,2,D]."""
    D = cfg["dim"]

    def mlp(pref_e, pref_p, v, chunks):
        emb = sinusoidal_embedding_1d(cfg["freq_dim"], v).float()
        e = F.linear(F.silu(F.linear(emb, sd[pref_e + ".0.weight"], sd[pref_e + ".0.bias"])), sd[pref_e + ".2.weight"],
                     sd[pref_e + ".2.bias"])
        e0 = F.linear(F.silu(e), sd[pref_p + ".1.weight"], sd[pref_p + ".1.bias"]).unflatten(1, (chunks, D))
        return e, e0

    e, e0 = mlp("time_embedding", "time_projection", t_flat, 6)
    de, de0 = mlp("density_embedding", "density_projection", density, 2)
    return e, e0, de, de0


def block_forward(sd, cfg, i, x, e0, dens0, grid, angles, ctx, policy, taps: Optional[dict] = None):
    """One WanAttentionBlock (:422-472) for ONE sample. x: [L, D] fp32; e0: [L, 6, D] or [6, D]; dens0: [2, D]."""
    D, nh, eps = cfg["dim"], cfg["num_heads"], cfg["eps"]
    hd = D // nh
    p = f"blocks.{i}."
    L = x.shape[0]
    m = sd[p + "modulation"][0] + e0                       # [.., 6, D]  :444-448
    m = [m[..., j, :] for j in range(6)]
    dm = sd[p + "modulation_density"][0] + dens0           # [2, D]     :449

    def ln(u):
        return F.layer_norm(u, (D,), eps=eps)

    # self-attention :452-456, :230-262
    h = _r(ln(x) * (1 + m[1]) + m[0] + dm[0], policy)
    sa = p + "self_attn."
    q = _rmsnorm(_linear(h, sd[sa + "q.weight"], sd[sa + "q.bias"], policy), sd[sa + "norm_q.weight"], eps, policy)
    k = _rmsnorm(_linear(h, sd[sa + "k.weight"], sd[sa + "k.bias"], policy), sd[sa + "norm_k.weight"], eps, policy)
    v = _linear(h, sd[sa + "v.weight"], sd[sa + "v.bias"], policy)
    q = _r(rope_apply(q.view(L, nh, hd), grid, angles), policy)
    k = _r(rope_apply(k.view(L, nh, hd), grid, angles), policy)
    a = _attention(q, k, v.view(L, nh, hd), policy).reshape(L, D)
    y = _linear(a, sd[sa + "o.weight"], sd[sa + "o.bias"], policy)
    x = x + y * m[2]
    if taps is not None:
        taps[f"b{i}.after_self"] = x.clone()

    # cross-attention :461, :353-371 — all text_len context slots attended, no mask (F8)
    ca = p + "cross_attn."
    c = F.layer_norm(x, (D,), sd[p + "norm3.weight"], sd[p + "norm3.bias"], eps=eps)
    q = _rmsnorm(_linear(c, sd[ca + "q.weight"], sd[ca + "q.bias"], policy), sd[ca + "norm_q.weight"], eps, policy)
    k = _rmsnorm(_linear(ctx, sd[ca + "k.weight"], sd[ca + "k.bias"], policy), sd[ca + "norm_k.weight"], eps, policy)
    v = _linear(ctx, sd[ca + "v.weight"], sd[ca + "v.bias"], policy)
    Lc = ctx.shape[0]
    a = _attention(q.view(L, nh, hd), k.view(Lc, nh, hd), v.view(Lc, nh, hd), policy).reshape(L, D)
    x = x + _linear(a, sd[ca + "o.weight"], sd[ca + "o.bias"], policy)
    if taps is not None:
        taps[f"b{i}.after_cross"] = x.clone()

    # ffn :464-468
    g = _r(ln(x) * (1 + m[4]) + m[3] + dm[1], policy)
    u = _r(_gelu_tanh(_linear(g, sd[p + "ffn.0.weight"], sd[p + "ffn.0.bias"], policy)), policy)
    y = _linear(u, sd[p + "ffn.2.weight"], sd[p + "ffn.2.bias"], policy)
    x = x + y * m[5]
    return x


def forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, t: torch.Tensor, context: List[torch.Tensor],
            seq_len: int, y: torch.Tensor, full_ref: torch.Tensor, additional_control: torch.Tensor,
            density: torch.Tensor, policy: str = "fp32", num_layers: Optional[int] = None,
            taps: Optional[dict] = None, teacache: Optional["TeaCacheOracle"] = None,
            cond_flag: bool = True) -> torch.Tensor:
    """Restates forward() :817-1123 for the FlexAM configuration (y, full_ref, additional_control, density given;
    clip_fea / y_camera / subject_ref absent; sp_world_size 1; TeaCache :977-1051 when `teacache` is given).
    Returns [B, out_dim, F, H, W] fp32."""
    D, C = cfg["dim"], cfg["out_dim"]
    B, _, Fr, H, W = x.shape
    Hp, Wp = H // 2, W // 2
    nl = cfg["num_layers"] if num_layers is None else num_layers
    angles = rope_angles(D // cfg["num_heads"])

    # control fuser + channel concat :868-883
    cnn_out = cnn_fuser(sd, y[:, :C], additional_control, policy)
    xin = torch.cat([x, cnn_out, y[:, C:]], dim=1)                       # [B, in_dim, F, H, W]

    # patch / ref embedding :885-899 (Conv3d k=s=(1,2,2) and Conv2d k=s=2), ref tokens PREPENDED
    pe = _r(F.conv3d(_r(xin, policy), sd["patch_embedding.weight"], sd["patch_embedding.bias"], stride=(1, 2, 2)), policy)
    tok = pe.flatten(2).transpose(1, 2)                                   # [B, L0, D], token order (f, h, w)
    rf = _r(F.conv2d(_r(full_ref, policy), sd["ref_conv.weight"], sd["ref_conv.bias"], stride=2), policy)
    rtok = rf.flatten(2).transpose(1, 2)                                  # [B, R, D]
    xs = torch.cat([rtok, tok], dim=1)
    R = rtok.shape[1]
    L = seq_len + R
    assert xs.shape[1] == L
    grid = (Fr + 1, Hp, Wp)

    # per-token timesteps are left-padded with t[:, -1] for the ref tokens :900-904
    per_token = t.dim() == 2
    if per_token and t.shape[1] < L:
        t = torch.cat([t[:, -1:].repeat(1, L - t.shape[1]), t], dim=1)
    e, e0, de, de0 = embed_mlps(sd, cfg, t.flatten().float(), density.float(), policy)
    if per_token:
        e, e0 = e.view(B, L, D), e0.view(B, L, 6, D)

    # context: zero-pad to text_len BEFORE the embedding MLP :958-964
    ctx = torch.stack([torch.cat([u, u.new_zeros(cfg["text_len"] - u.shape[0], u.shape[1])]) for u in context])
    ctx = _linear(_r(_gelu_tanh(_linear(ctx, sd["text_embedding.0.weight"], sd["text_embedding.0.bias"], policy)), policy),
                  sd["text_embedding.2.weight"], sd["text_embedding.2.bias"], policy)
    if taps is not None:
        taps["x0"], taps["e"], taps["e0"], taps["ctx"], taps["cnn_out"] = xs.clone(), e, e0, ctx, cnn_out

    # TeaCache :977-1051: decide from the modulated timestep embedding of the LAST token, then either run the block
    # stack (remembering x_after - x_before) or re-apply the remembered residual of the last len(x) samples
    run_blocks = True
    if teacache is not None:
        run_blocks = teacache.decide(e0[:, -1, :] if per_token else e0, cond_flag)
    if run_blocks:
        after = []
        for b in range(B):
            xb = xs[b]
            for i in range(nl):
                xb = block_forward(sd, cfg, i, xb, e0[b], de0[b], grid, angles, ctx[b], policy,
                                   taps if (taps is not None and b == 0) else None)
            after.append(xb)
        after = torch.stack(after)
        if teacache is not None:
            teacache.store(after - xs, cond_flag)
    else:
        after = xs + teacache.residual(cond_flag)[-B:]
    if taps is not None:
        taps["x_final"] = after[0].clone()

    outs = []
    for b in range(B):
        xb = after[b]
        # head :493-507 — uses e (not e0) and a single density chunk
        hm = sd["head.modulation"][0] + e[b].unsqueeze(-2)               # [.., 2, D]
        hd = sd["head.modulation_density"][0, 0] + de[b]
        z = F.layer_norm(xb, (D,), eps=cfg["eps"]) * (1 + hm[..., 1, :]) + hm[..., 0, :] + hd
        o = _linear(z, sd["head.head.weight"], sd["head.head.bias"], policy)      # [L, 4*C]
        o = o[R:]                                                         # strip ref tokens from the FRONT :1106-1109
        # unpatchify :1142-1148  'fhwpqrc->cfphqwr' with p == 1
        o = o.view(Fr, Hp, Wp, 1, 2, 2, C)
        o = torch.einsum("fhwpqrc->cfphqwr", o).reshape(C, Fr, H, W)
        outs.append(o)
    if teacache is not None and cond_flag:                               # :1119-1122
        teacache.cnt += 1
        if teacache.cnt == teacache.num_steps:
            teacache.reset()
    return torch.stack(outs)


def forward_cfg_skip(sd, cfg, x, t, context, seq_len, y, full_ref, additional_control, density, cfg_skip_ratio,
                     current_step: int, num_steps: int, **kw) -> torch.Tensor:
    """forward() under its @cfg_skip() wrapper (FlexAM/utils/cfg_optimization.py:5-38): late in sampling only the
    second (conditional) half of the batch is evaluated and the result is duplicated."""
    bs = len(x)
    skip = bs >= 2 and cfg_skip_ratio is not None and current_step >= num_steps * (1 - cfg_skip_ratio)
    if skip:
        h = bs // 2
        x, t, context, y, full_ref, additional_control, density = (
            u[h:] for u in (x, t, context, y, full_ref, additional_control, density))
    out = forward(sd, cfg, x, t, context, seq_len, y, full_ref, additional_control, density, **kw)
    return torch.cat([out, out], dim=0) if skip else out


class TeaCacheOracle:
    """Restates FlexAM/models/cache_utils.py:21-76 (state) and the decision block of forward :978-1000."""

    def __init__(self, coefficients, num_steps: int, rel_l1_thresh: float = 0.0, num_skip_start_steps: int = 0):
        self.coefficients = [float(c) for c in coefficients]
        self.num_steps, self.rel_l1_thresh, self.num_skip_start_steps = num_steps, rel_l1_thresh, num_skip_start_steps
        self.decisions = []          # test instrumentation: one bool per cond_flag forward
        self.reset()

    def reset(self):
        self.cnt, self.should_calc, self.acc = 0, True, 0.0
        self.prev_mod = self.res_cond = self.res_uncond = None

    def rescale(self, d: float) -> float:      # np.poly1d(coefficients)(d): highest power first
        y = 0.0
        for c in self.coefficients:
            y = y * d + c
        return y

    def decide(self, modulated_inp: torch.Tensor, cond_flag: bool) -> bool:
        if not cond_flag:
            return self.should_calc
        if self.cnt < self.num_skip_start_steps:
            self.should_calc, self.acc = True, 0.0
        else:
            d = ((modulated_inp - self.prev_mod).abs().mean() / self.prev_mod.abs().mean()).item()
            self.acc += self.rescale(d)
            if self.acc < self.rel_l1_thresh:
                self.should_calc = False
            else:
                self.should_calc, self.acc = True, 0.0
        self.prev_mod = modulated_inp.clone()
        self.decisions.append(self.should_calc)
        return self.should_calc

    def store(self, residual: torch.Tensor, cond_flag: bool):
        if cond_flag:
            self.res_cond = residual
        else:
            self.res_uncond = residual

    def residual(self, cond_flag: bool) -> torch.Tensor:
        return self.res_cond if cond_flag else self.res_uncond


def to_torch_sd(np_sd: dict, device="cpu") -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(v).to(device) for k, v in np_sd.items()}


def flops_per_cfg_branch(cfg: dict, L0: int, R: int, n_pixels: int) -> float:
    """Algorithmic FLOPs of one forward for ONE sample (SURVEY.md §8d convention)."""
    D, Fd, Lc = cfg["dim"], cfg["ffn_dim"], cfg["text_len"]
    L = L0 + R
    per_layer = 2 * L * D * D * 4 + 4 * L * L * D + 2 * L * D * D * 2 + 2 * Lc * D * D * 2 + 4 * L * Lc * D + \
        2 * L * D * Fd * 2
    front = 2 * L0 * D * cfg["in_dim"] * 4 + 2 * R * D * cfg["out_dim"] * 4 + 2 * L * D * cfg["out_dim"] * 4 + \
        2 * Lc * (cfg["text_dim"] * D + D * D)
    cnn = 2 * n_pixels * 9 * (cfg["in_dim_cnn"] * 192 + 192 * 192 + 192 * 96 + 96 * 96) + 2 * n_pixels * 96 * cfg["out_dim_cnn"]
    return float(cfg["num_layers"] * per_layer + front + cnn)
