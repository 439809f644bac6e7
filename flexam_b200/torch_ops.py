"""``torch.library`` registration of the block-level operators (SURVEY.md §8b "Binding").

``flexam_b200.ops`` is the binding the engine itself uses (plain Python functions over ctypes). This module exposes the
same launches as dispatcher operators ``torch.ops.flexam_b200.*`` so that code traced by ``torch.compile`` /
``make_fx`` / ``FakeTensorMode`` — the reference compiles its ``WanAttentionBlock`` modules one by one in the ComfyUI
node (comfyui/comfyui_nodes.py:69-73) — sees opaque, correctly typed calls instead of ctypes:

* every operator writes into a caller-provided ``out`` / in-place tensor (``mutates_args``) and returns nothing, like
  the C ABI underneath (no allocation inside), so no shape function is needed: the fake implementation is a no-op;
* they are registered for the CUDA dispatch key ONLY. A CPU tensor raises ``NotImplementedError`` from the
  dispatcher: there is no torch/CPU fallback behind these names either.

Nothing on the product path depends on this module; importing it does not load the shared library.
"""
from __future__ import annotations

from typing import List, Optional

from torch import Tensor
from torch.library import custom_op

from . import ops

NAMESPACE = "flexam_b200"


@custom_op(f"{NAMESPACE}::gemm", mutates_args=("out",), device_types="cuda")
def gemm(a: Tensor, w: Tensor, bias: Optional[Tensor], out: Tensor, epilogue: int, gate_mod: Optional[Tensor] = None,
         gate_e: Optional[Tensor] = None, row_idx: Optional[Tensor] = None) -> None:
    """fx_gemm_bf16: out = epilogue(a @ w^T + bias); FX_EPI_RESID_F32 accumulates into ``out`` (the residual stream).
    Replaces the nn.Linear call sites wan_transformer3d_FlexAM.py:242-261, :363-370, :414-416, :456-468, :506."""
    ops.gemm(a, w, bias, out, epilogue, gate_mod=gate_mod, gate_e=gate_e, row_idx=row_idx)


@custom_op(f"{NAMESPACE}::fmha", mutates_args=("out",), device_types="cuda")
def fmha(q: Tensor, k: Tensor, v: Tensor, out: Tensor, scale: float) -> None:
    """fx_fmha_fwd: non-causal attention over [B, L, H, 128] views (attention_utils.py:174-233; :251-256, :367)."""
    ops.fmha(q, k, v, out, scale)


@custom_op(f"{NAMESPACE}::ln_modulate", mutates_args=("out",), device_types="cuda")
def ln_modulate(x: Tensor, out: Tensor, eps: float, shift_mod: Tensor, scale_mod: Tensor, shift_e: Tensor,
                scale_e: Tensor, e_stride: int, row_idx: Optional[Tensor], dens_mod: Optional[Tensor],
                dens: Optional[Tensor], dens_stride: int, rows_per_batch: int) -> None:
    """fx_ln_modulate: LayerNorm + adaLN modulation + density shift, fp32 -> bf16 (:444-453, :464-465, :493-507)."""
    ops.ln_modulate(x, out, eps, shift_mod, scale_mod, shift_e, scale_e, e_stride, row_idx, dens_mod, dens,
                    dens_stride, rows_per_batch)


@custom_op(f"{NAMESPACE}::ln_scale_shift", mutates_args=("out",), device_types="cuda")
def ln_scale_shift(x: Tensor, out: Tensor, eps: float, scale: Tensor, shift: Tensor, row_stride: int,
                   row_idx: Optional[Tensor]) -> None:
    """fx_ln_scale_shift: the block-level LayerNorm sites with the modulation rows combined per (timestep, sample)."""
    ops.ln_scale_shift(x, out, eps, scale, shift, row_stride, row_idx)


@custom_op(f"{NAMESPACE}::ln_affine", mutates_args=("out",), device_types="cuda")
def ln_affine(x: Tensor, out: Tensor, eps: float, gamma: Tensor, beta: Tensor) -> None:
    """fx_ln_affine: norm3 (:405-407, :461)."""
    ops.ln_affine(x, out, eps, gamma, beta)


@custom_op(f"{NAMESPACE}::rmsnorm_rope", mutates_args=("x",), device_types="cuda")
def rmsnorm_rope(x: Tensor, weight: Tensor, eps: float, freqs: Optional[Tensor], grid: List[int], tok_offset: int,
                 rows_per_batch: int, weight2: Optional[Tensor]) -> None:
    """fx_rmsnorm_rope, in place: WanRMSNorm (:173-189) + rope_apply (:135-170) on q (and k with ``weight2``)."""
    ops.rmsnorm_rope(x, weight, eps, freqs, tuple(grid), tok_offset, rows_per_batch, weight2)


@custom_op(f"{NAMESPACE}::modulation_tables", mutates_args=("tab",), device_types="cuda")
def modulation_tables(mod: Tensor, dmod: Tensor, e0: Tensor, de0: Tensor, tab: Tensor) -> None:
    """fx_modulation_tables: (modulation + e0).chunk(6) combined once per (timestep, sample) (:444-449)."""
    ops.modulation_tables(mod, dmod, e0, de0, tab)


@custom_op(f"{NAMESPACE}::unpatchify", mutates_args=("out",), device_types="cuda")
def unpatchify(head: Tensor, out: Tensor) -> None:
    """fx_unpatchify (:1126-1149)."""
    ops.unpatchify(head, out)


@custom_op(f"{NAMESPACE}::cfg_euler_step", mutates_args=("lat",), device_types="cuda")
def cfg_euler_step(vu: Tensor, vc: Tensor, guidance: float, dsigma: float, lat: Tensor, mask: Optional[Tensor],
                   pinned: Optional[Tensor]) -> None:
    """fx_cfg_euler_step: CFG combine + Euler step + first-frame re-pin (pipeline…:926-934)."""
    ops.cfg_euler_step(vu, vc, guidance, dsigma, lat, mask, pinned)


OPERATORS = ("gemm", "fmha", "ln_modulate", "ln_scale_shift", "ln_affine", "rmsnorm_rope", "modulation_tables",
             "unpatchify", "cfg_euler_step")
