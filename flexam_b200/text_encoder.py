"""umT5 text encoder on the native kernels.

``WanT5EncoderModel`` mirrors the reference class (FlexAM/models/wan_text_encoder.py:256-304, cited as :line): same
constructor kwargs, same parameter tree / state_dict keys (``token_embedding.weight``, ``blocks.i.{norm1, attn.{q,k,v,o},
norm2, ffn.{gate.0, fc1, fc2}, pos_embedding.embedding}.weight``, ``norm.weight``), same ``forward(input_ids,
attention_mask) -> (hidden_states,)`` — what ``Wan2_2FunControlPipeline_FlexAM._get_t5_prompt_embeds`` calls once per
prompt batch. The forward is a fixed sequence of ``fx_*`` launches: packed q|k|v and the gate / fc1 / fc2 / o
projections on the tcgen05 GEMM, T5LayerNorm, the biased softmax attention (head_dim 64, per-layer relative-position
table), the bf16 residual adds and the gated GELU as row kernels (csrc/textenc.cu). No torch compute, no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .lib import FX_EPI_BF16, FlexamNativeError

bf16, i32 = torch.bfloat16, torch.int32


def relative_position_bucket(lq: int, lk: int, num_buckets: int, max_dist: int = 128) -> torch.Tensor:
    """T5RelativeEmbedding._relative_position_bucket (:235-253), bidirectional: [lq, lk] bucket of (key - query)."""
    rel = torch.arange(lk).unsqueeze(0) - torch.arange(lq).unsqueeze(1)
    nb = num_buckets // 2
    buckets = (rel > 0).long() * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = max_exact + (torch.log(rel.float() / max_exact) / math.log(max_dist / max_exact) * (nb - max_exact)).long()
    large = torch.min(large, torch.full_like(large, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


class T5Engine:
    """Packed weights + workspaces + the launch sequence for a parameter mapping with the reference's keys."""

    def __init__(self, params: Dict[str, torch.Tensor], cfg: dict, device: torch.device):
        self.cfg, self.device, self.params = dict(cfg), torch.device(device), params
        self.H = cfg["num_heads"]
        if cfg["dim_attn"] // self.H != 64:
            raise FlexamNativeError("native T5 attention supports head_dim 64 (umT5-XXL: 4096 / 64 heads)")
        self._ws = {}
        self._bias = {}
        self.launches = 0
        self._pack()

    def _pack(self):
        P = self.params
        for k, v in P.items():
            if v.device != self.device or v.dtype != bf16:
                raise FlexamNativeError(f"parameter {k}: expected bf16 on {self.device}, got {v.dtype} on {v.device}")
        self.blk = []
        for i in range(self.cfg["num_layers"]):
            p = f"blocks.{i}."
            self.blk.append({
                "wqkv": torch.cat([P[p + "attn.q.weight"], P[p + "attn.k.weight"], P[p + "attn.v.weight"]], 0).contiguous(),
                "wo": P[p + "attn.o.weight"], "n1": P[p + "norm1.weight"], "n2": P[p + "norm2.weight"],
                # gate | fc1 as ONE projection (N = 2 * dim_ffn): 4.3 waves of 256x256 tiles instead of 2 x 2.2 at M = 1024
                "wgu": torch.cat([P[p + "ffn.gate.0.weight"], P[p + "ffn.fc1.weight"]], 0).contiguous(),
                "w2": P[p + "ffn.fc2.weight"],
                "pos": P[p + "pos_embedding.embedding.weight"],
            })
        self._versions = tuple(p._version for p in P.values())
        self._bias.clear()

    def _buf(self, name, shape, dtype):
        key = (name, tuple(shape), dtype)
        t = self._ws.get(key)
        if t is None:
            t = torch.empty(tuple(shape), dtype=dtype, device=self.device)
            self._ws[key] = t
        return t

    def _bias_tables(self, L: int):
        """bias_rel[layer] bf16 [H, 2L-1]: pos_embedding[bucket(d)][h] for d = j - i in [-(L-1), L-1] (:219-232).
        Index glue on a [num_buckets, H] weight, once per (L, weight version)."""
        t = self._bias.get(L)
        if t is None:
            bucket = relative_position_bucket(L, L, self.cfg["num_buckets"])
            d = torch.cat([bucket[L - 1, :L - 1], bucket[0]]).to(self.device)      # rel = -(L-1)..-1 then 0..L-1
            t = [w["pos"][d].t().contiguous() for w in self.blk]
            self._bias[L] = t
        return t

    @torch.no_grad()
    def forward(self, input_ids: torch.Tensor, attention_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if tuple(p._version for p in self.params.values()) != self._versions:
            self._pack()
        cfg, dev = self.cfg, self.device
        D, A, Fd, H = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"], self.H
        ids = input_ids.to(dev, torch.int64).contiguous()
        B, L = ids.shape
        M = B * L
        mask = None if attention_mask is None else attention_mask.to(dev).ne(0).to(i32).contiguous()
        bias = self._bias_tables(L)
        self.launches = 0
        x = torch.empty((M, D), dtype=bf16, device=dev)
        ops.embedding(ids, self.params["token_embedding.weight"], x)
        h = self._buf("h", (M, D), bf16)
        qkv = self._buf("qkv", (M, 3 * A), bf16)
        att = self._buf("att", (M, A), bf16)
        y = self._buf("y", (M, D), bf16)
        gu = self._buf("gate_fc1", (M, 2 * Fd), bf16)
        u = self._buf("ffn", (M, Fd), bf16)
        for i, w in enumerate(self.blk):
            ops.t5_layernorm(x, w["n1"], h)
            ops.gemm(h, w["wqkv"], None, qkv, FX_EPI_BF16)
            ops.t5_attention(qkv, bias[i], mask, att, B, L, H)
            ops.gemm(att, w["wo"], None, y, FX_EPI_BF16)
            ops.add_bf16_(x, y)                                      # x = x + attn(norm1(x))            :161
            ops.t5_layernorm(x, w["n2"], h)
            ops.gemm(h, w["wgu"], None, gu, FX_EPI_BF16)
            ops.gated_gelu(gu[:, Fd:], gu[:, :Fd], u)                # fc1(x) * GELU(gate(x))            :126
            ops.gemm(u, w["w2"], None, y, FX_EPI_BF16)
            ops.add_bf16_(x, y)                                      # x = x + ffn(norm2(x))             :162
            self.launches += 10
        out = torch.empty((M, D), dtype=bf16, device=dev)
        ops.t5_layernorm(x, self.params["norm.weight"], out)
        self.launches += 2
        return out.view(B, L, D)


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


def param_shapes(cfg: dict) -> Dict[str, tuple]:
    D, A, Fd = cfg["dim"], cfg["dim_attn"], cfg["dim_ffn"]
    out = {"token_embedding.weight": (cfg["vocab"], D)}
    for i in range(cfg["num_layers"]):
        p = f"blocks.{i}."
        out.update({p + "norm1.weight": (D,), p + "attn.q.weight": (A, D), p + "attn.k.weight": (A, D),
                    p + "attn.v.weight": (A, D), p + "attn.o.weight": (D, A), p + "norm2.weight": (D,),
                    p + "ffn.gate.0.weight": (Fd, D), p + "ffn.fc1.weight": (Fd, D), p + "ffn.fc2.weight": (D, Fd),
                    p + "pos_embedding.embedding.weight": (cfg["num_buckets"], cfg["num_heads"])})
    out["norm.weight"] = (D,)
    return out


class WanT5EncoderModel(nn.Module):
    """Drop-in for FlexAM.models.WanT5EncoderModel (:256-304) with a native forward (shared_pos=False, dropout 0 as
    in config/wan2.2/wan_civitai_5b_FlexAM.yaml:20-32)."""

    def __init__(self, vocab, dim, dim_attn, dim_ffn, num_heads, num_layers, num_buckets, shared_pos=False, dropout=0.0,
                 dtype=torch.bfloat16, device=None):
        super().__init__()
        if shared_pos:
            raise FlexamNativeError("native umT5 path: per-layer relative-position tables (shared_pos=False) only")
        self.config = dict(vocab=vocab, dim=dim, dim_attn=dim_attn, dim_ffn=dim_ffn, num_heads=num_heads,
                           num_layers=num_layers, num_buckets=num_buckets)
        self.dim, self.dim_attn, self.dim_ffn = dim, dim_attn, dim_ffn
        self.num_heads, self.num_layers, self.num_buckets, self.shared_pos = num_heads, num_layers, num_buckets, shared_pos
        self.blocks = nn.ModuleList([nn.Module() for _ in range(num_layers)])
        for name, shape in param_shapes(self.config).items():
            _set_param(self, name, nn.Parameter(torch.empty(shape, dtype=dtype, device=device), requires_grad=False))
        self._engine: Optional[T5Engine] = None

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @classmethod
    def from_pretrained(cls, pretrained_model_path, additional_kwargs={}, low_cpu_mem_usage=False,
                        torch_dtype=torch.bfloat16):
        """The reference loader's contract (:306-393): constructor kwargs filtered from ``additional_kwargs``
        (config/wan2.2/wan_civitai_5b_FlexAM.yaml:20-32), a ``.safetensors`` or ``torch.load`` checkpoint, non-strict load
        with the missing / unexpected keys printed, cast to ``torch_dtype``. ``low_cpu_mem_usage`` is accepted and gives
        the same result (the reference's meta-device path is an optimisation of the same load)."""
        import inspect
        valid = set(inspect.signature(cls.__init__).parameters) - {"self"}
        model = cls(**{k: v for k, v in dict(additional_kwargs).items() if k in valid})
        if str(pretrained_model_path).endswith(".safetensors"):
            from safetensors.torch import load_file
            state_dict = load_file(pretrained_model_path)
        else:
            state_dict = torch.load(pretrained_model_path, map_location="cpu")
        own = model.state_dict()
        m, u = model.load_state_dict({k: v.to(own[k].dtype) if k in own else v for k, v in state_dict.items()}, strict=False)
        print(f"### missing keys: {len(m)}; \n### unexpected keys: {len(u)};")
        print(m, u)
        return model.to(torch_dtype)

    def engine(self) -> T5Engine:
        params = {k: v.detach() for k, v in self.named_parameters()}
        if self._engine is None or tuple(p.data_ptr() for p in self._engine.params.values()) != \
                tuple(p.data_ptr() for p in params.values()):
            self._engine = T5Engine(params, self.config, next(iter(params.values())).device)
        return self._engine

    def forward(self, input_ids: Optional[torch.LongTensor] = None, attention_mask: Optional[torch.Tensor] = None):
        eng = self.engine()
        guard = torch.cuda.device(eng.device) if eng.device.type == "cuda" else _null()
        with guard, ops.stream_scope():
            return (eng.forward(input_ids, attention_mask),)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
