"""TEST INFRASTRUCTURE — CPU restatement of the reference umT5 text encoder.

Restates ``WanT5EncoderModel.forward`` (reference file ``FlexAM/models/wan_text_encoder.py``, cited as :line) — token
embedding, 24 x {T5LayerNorm, self-attention with a per-layer relative-position bias and no 1/sqrt(d) scaling, T5LayerNorm,
gated-GELU feed-forward}, final T5LayerNorm — as one functional pass over a ``{state_dict key: tensor}`` mapping. Only
``tests/`` and ``bench.py`` may import it; ``flexam_b200`` never does.

Pinning: the reference has no tests or vectors for this module either; ``oracle/make_golden.py t5_tiny`` runs the REAL
module (``oracle/ref_import.build_reference_t5``) on ``oracle/synth`` weights and stores its fp32 CPU output under
``tests/golden/t5_tiny.npz``; ``tests/test_t5.py`` checks this file against it and, when /root/reference is mounted,
against the live module.

Precision policies: "fp32" (what the module computes on CPU in fp32) and "bf16" (the module as the pipeline runs it,
``torch_dtype=bf16`` weights AND activations: every torch op rounds its result to bf16 — the residual adds, the two
roundings inside T5LayerNorm :51-56, the score einsum, the bias add, the softmax output, GELU's elementwise chain).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F


def _r(t, policy):
    return t.to(torch.bfloat16).to(torch.float32) if policy == "bf16" else t


def relative_position_bucket(lq: int, lk: int, num_buckets: int = 32, max_dist: int = 128) -> torch.Tensor:
    """:235-253 for the bidirectional case: [lq, lk] int64 bucket ids of (key position - query position)."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    nb = num_buckets // 2
    buckets = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


def t5_layer_norm(x, w, eps, policy):
    # :51-56 — x * rsqrt(mean(x.float()^2) + eps) is an fp32 product, cast to the weight dtype, then times the weight
    y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + eps)
    return _r(w * _r(y, policy), policy)


def gelu_chain(x, policy):
    # :38-41 — written with elementwise torch ops; in bf16 every intermediate is rounded
    r = lambda t: _r(t, policy)  # noqa: E731
    inner = r(x + r(0.044715 * r(torch.pow(x, 3.0))))
    return r(r(0.5 * x) * r(1.0 + r(torch.tanh(r(math.sqrt(2.0 / math.pi) * inner)))))


def forward(sd: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor, attention_mask: torch.Tensor = None,
            policy: str = "fp32") -> torch.Tensor:
    """:291-304. input_ids [B, L] int64, attention_mask [B, L] (1 = real token) or None -> [B, L, dim] fp32."""
    n, eps = cfg["num_heads"], 1e-6
    c = cfg["dim_attn"] // n
    B, L = input_ids.shape
    x = _r(sd["token_embedding.weight"][input_ids], policy)
    buckets = relative_position_bucket(L, L, cfg["num_buckets"]).to(x.device)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        h = t5_layer_norm(x, sd[p + "norm1.weight"], eps, policy)
        q = _r(F.linear(h, sd[p + "attn.q.weight"]), policy).view(B, L, n, c)
        k = _r(F.linear(h, sd[p + "attn.k.weight"]), policy).view(B, L, n, c)
        v = _r(F.linear(h, sd[p + "attn.v.weight"]), policy).view(B, L, n, c)
        # :95-103 attn_bias = zeros + pos_bias (+ mask fill with finfo.min); per-layer embedding when shared_pos is False
        bias = _r(sd[p + "pos_embedding.embedding.weight"], policy)[buckets].permute(2, 0, 1).unsqueeze(0)   # [1,n,L,L]
        bias = bias.expand(B, n, L, L).clone()
        if attention_mask is not None:
            fill = torch.finfo(torch.bfloat16 if policy == "bf16" else torch.float32).min
            bias.masked_fill_(attention_mask.view(B, 1, 1, L) == 0, fill)
        s = _r(_r(torch.einsum("binc,bjnc->bnij", q, k), policy) + bias, policy)        # :106 (no scaling)
        a = _r(F.softmax(s.float(), dim=-1), policy)                                      # :107
        o = _r(torch.einsum("bnij,bjnc->binc", a, v), policy).reshape(B, L, n * c)
        x = _r(x + _r(F.linear(o, sd[p + "attn.o.weight"]), policy), policy)             # :161
        h = t5_layer_norm(x, sd[p + "norm2.weight"], eps, policy)
        g = gelu_chain(_r(F.linear(h, sd[p + "ffn.gate.0.weight"]), policy), policy)
        u = _r(_r(F.linear(h, sd[p + "ffn.fc1.weight"]), policy) * g, policy)            # :126
        x = _r(x + _r(F.linear(u, sd[p + "ffn.fc2.weight"]), policy), policy)            # :162
    return t5_layer_norm(x, sd["norm.weight"], eps, policy)


def forward_library(sd: Dict[str, torch.Tensor], cfg: dict, input_ids: torch.Tensor,
                    attention_mask: torch.Tensor = None) -> torch.Tensor:
    """The same forward with stock torch ops on tensors of the weights' dtype — what the reference module executes when
    the pipeline runs it in bf16 on a GPU (cuBLAS Linear, bf16 einsum, fp32 softmax): bench.py's library-path leg for
    the text encoder, and the real bf16 numerics of those libraries as a second parity anchor."""
    n = cfg["num_heads"]
    c = cfg["dim_attn"] // n
    B, L = input_ids.shape

    def norm(x, w):                                            # :51-56
        y = x * torch.rsqrt(x.float().pow(2).mean(dim=-1, keepdim=True) + 1e-6)
        return w * y.type_as(w)

    def gelu(x):                                               # :38-41
        return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))
    x = sd["token_embedding.weight"][input_ids]
    buckets = relative_position_bucket(L, L, cfg["num_buckets"]).to(x.device)
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        h = norm(x, sd[p + "norm1.weight"])
        q = F.linear(h, sd[p + "attn.q.weight"]).view(B, L, n, c)
        k = F.linear(h, sd[p + "attn.k.weight"]).view(B, L, n, c)
        v = F.linear(h, sd[p + "attn.v.weight"]).view(B, L, n, c)
        bias = x.new_zeros(B, n, L, L)
        bias += sd[p + "pos_embedding.embedding.weight"][buckets].permute(2, 0, 1).unsqueeze(0)
        if attention_mask is not None:
            bias.masked_fill_(attention_mask.view(B, 1, 1, L) == 0, torch.finfo(x.dtype).min)
        a = torch.einsum("binc,bjnc->bnij", q, k) + bias
        a = F.softmax(a.float(), dim=-1).type_as(a)
        o = torch.einsum("bnij,bjnc->binc", a, v).reshape(B, L, n * c)
        x = x + F.linear(o, sd[p + "attn.o.weight"])
        h = norm(x, sd[p + "norm2.weight"])
        x = x + F.linear(F.linear(h, sd[p + "ffn.fc1.weight"]) * gelu(F.linear(h, sd[p + "ffn.gate.0.weight"])),
                         sd[p + "ffn.fc2.weight"])
    return norm(x, sd["norm.weight"])


T5_CONFIGS = {
    # umT5-XXL encoder as the FlexAM yaml builds it (config/wan2.2/wan_civitai_5b_FlexAM.yaml:20-32)
    "real": dict(vocab=256384, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=24, num_buckets=32),
    # real width / head layout, 2 layers, small vocabulary: CPU-runnable in seconds
    "real2": dict(vocab=1024, dim=4096, dim_attn=4096, dim_ffn=10240, num_heads=64, num_layers=2, num_buckets=32),
    "tiny": dict(vocab=300, dim=256, dim_attn=256, dim_ffn=512, num_heads=4, num_layers=2, num_buckets=32),
}


def param_specs(cfg: dict):
    """[(state_dict key, shape, std, mean)] of WanT5EncoderModel(shared_pos=False) (:256-289)."""
    D, A, Fd = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"]
    specs = [("token_embedding.weight", (cfg["vocab"], D), 1.0, 0.0)]
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        specs += [(p + "norm1.weight", (D,), 0.1, 1.0), (p + "attn.q.weight", (A, D), 0.2 * D ** -0.5, 0.0),   # no 1/sqrt(d) in T5: logits of std ~1.6
                  (p + "attn.k.weight", (A, D), D ** -0.5, 0.0), (p + "attn.v.weight", (A, D), D ** -0.5, 0.0),
                  (p + "attn.o.weight", (D, A), A ** -0.5, 0.0), (p + "norm2.weight", (D,), 0.1, 1.0),
                  (p + "ffn.gate.0.weight", (Fd, D), D ** -0.5, 0.0), (p + "ffn.fc1.weight", (Fd, D), D ** -0.5, 0.0),
                  (p + "ffn.fc2.weight", (D, Fd), Fd ** -0.5, 0.0),
                  (p + "pos_embedding.embedding.weight", (cfg["num_buckets"], cfg["num_heads"]), 0.5, 0.0)]
    specs.append(("norm.weight", (D,), 0.1, 1.0))
    return specs


def state_dict(cfg: dict, tag: str = "t5"):
    from oracle import synth
    return {name: synth.tensor(f"{tag}/{name}", shape, std, mean) for name, shape, std, mean in param_specs(cfg)}


def state_dict_torch(cfg: dict, device, dtype=None, tag: str = "t5"):
    from oracle import synth
    return {name: synth.tensor_torch(f"{tag}/{name}", shape, std, mean, device=device, dtype=dtype)
            for name, shape, std, mean in param_specs(cfg)}


def inputs(cfg: dict, L: int = 512, lens=(37, 120), tag: str = "t5in"):
    """Token ids of two prompts padded to L with id 0 and the tokenizer's attention mask (pipeline :203-230)."""
    import numpy as np
    from oracle import synth
    B = len(lens)
    ids = np.zeros((B, L), np.int64)
    mask = np.zeros((B, L), np.int64)
    for b, n in enumerate(lens):
        u = synth.tensor(f"{tag}/ids{b}", (n,), 1.0, 0.0, bf16=False)
        ids[b, :n] = 1 + (np.abs(u) / np.sqrt(3.0) * (cfg["vocab"] - 2)).astype(np.int64) % (cfg["vocab"] - 1)
        mask[b, :n] = 1
    return ids, mask
