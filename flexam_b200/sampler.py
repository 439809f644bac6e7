"""The denoising loop around the transformer (BASELINE config 4), B200-side.

Mirrors the control flow of ``Wan2_2FunControlPipeline_FlexAM.__call__`` step 7
(FlexAM/pipeline/pipeline_wan2_2_fun_control_FlexAM.py:843-934) for the FlexAM 5B configuration (mask video given,
VAE spatial ratio 16 => per-token timesteps and first-frame re-pinning, no camera branch, single transformer): per
step it builds the batch-of-2 ``[uncond, cond]`` call exactly as the pipeline does, runs the native transformer, and
replaces the pipeline's eight elementwise torch ops after it (chunk / CFG combine :926-928, ``scheduler.step``
:931, re-pin :933-934) with ONE ``fx_cfg_euler_step`` launch on an fp32 master copy of the latents whose values
stay bf16-representable (the pipeline keeps ``latents`` in the bf16 weight dtype).

The reference pipeline itself is unchanged by the drop-in; this module is the same loop for hosts that want the whole
sampling loop on the device without per-step host syncs (the pipeline's ``mask[:, :, 0].any()`` at :933 syncs every
step; here it is evaluated once). The default scheduler is diffusers' ``FlowMatchEulerDiscreteScheduler`` (absent from
the reference tree; ``requirements.txt:27`` pins ``diffusers>=0.30.1``); ``flow_match_euler_schedule`` restates its
``__init__`` + ``set_timesteps`` for ``use_dynamic_shifting=False`` (config/wan2.2/wan_civitai_5b_FlexAM.yaml:34-42).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .lib import FlexamNativeError

bf16, f32 = torch.bfloat16, torch.float32


def flow_match_euler_schedule(num_inference_steps: int, shift: float = 5.0,
                              num_train_timesteps: int = 1000) -> Tuple[np.ndarray, np.ndarray]:
    """(timesteps[n], sigmas[n+1]) of FlowMatchEulerDiscreteScheduler with a static shift: the constructor shifts the
    training sigmas once (which fixes sigma_max / sigma_min), ``set_timesteps`` spaces timesteps linearly between
    them and applies the shift again; a terminal sigma of 0 is appended. float32 like the scheduler's tensors."""
    n_train = float(num_train_timesteps)
    train = np.linspace(1.0, n_train, num_train_timesteps, dtype=np.float32)[::-1] / np.float32(n_train)
    train = shift * train / (1 + (shift - 1) * train)
    sigma_max, sigma_min = float(train[0]), float(train[-1])
    ts = np.linspace(sigma_max * n_train, sigma_min * n_train, num_inference_steps)
    s = ts / n_train
    s = (shift * s / (1 + (shift - 1) * s)).astype(np.float32)
    return (s * np.float32(n_train)).astype(np.float32), np.concatenate([s, np.zeros(1, np.float32)])


class DenoiseLoop:
    """Device-resident state of one sampling run: fp32 master latents, the pinned first-frame latents, the latent
    mask and the step-invariant control tensors (batch-of-2 copies built once, so the transformer's static cache —
    CNN control fuser, text embedding, cross K/V — hits on every step after the first)."""

    def __init__(self, transformer, latents: torch.Tensor, mask: torch.Tensor, masked_video_latents: torch.Tensor,
                 mask_latents: torch.Tensor, control_video_latents: torch.Tensor,
                 additional_control_latents: torch.Tensor, ref_image_latents: torch.Tensor,
                 negative_prompt_embeds: Sequence[torch.Tensor], prompt_embeds: Sequence[torch.Tensor],
                 density: float, guidance_scale: float = 6.0):
        if latents.dim() != 5 or latents.shape[0] != 1:
            raise FlexamNativeError("DenoiseLoop: latents must be [1, C, F, H, W] (one video per loop, CFG batch 2)")
        self.tf = transformer
        self.guidance = float(guidance_scale)
        dev = latents.device
        self.lat = latents.to(bf16).to(f32).contiguous()                     # master copy, bf16-representable values
        _, C, F, H, W = latents.shape
        m = mask.to(dev, f32)
        if m.shape != (1, 1, F, H, W):
            raise FlexamNativeError(f"DenoiseLoop: mask must be [1,1,{F},{H},{W}], got {tuple(m.shape)}")
        # pipeline :933 re-pins only when the whole first latent frame is masked out; decided ONCE, not per step
        self.repin = not bool(m[:, :, 0].any().item())
        self.mask_full = m.expand(1, C, F, H, W).contiguous() if self.repin else None
        self.pinned = masked_video_latents.to(dev, bf16).contiguous()
        # :892 per-token timestep factor; bf16 like the pipeline's mask, so mask * t rounds t to bf16 as it does there
        self.tok_mask = m[0, 0, :, ::2, ::2].reshape(-1).to(bf16).contiguous()
        self.seq_len = F * (H // 2) * (W // 2)
        if self.repin:   # the pipeline pins before the loop too (:688-690)
            self.lat = (((1 - self.mask_full) * self.pinned.float()).to(bf16).float()
                        + (self.mask_full * self.lat).to(bf16).float()).to(bf16).to(f32).contiguous()

        def two(u):
            return torch.cat([u, u]).to(dev, bf16).contiguous()
        # :857-877 control_latents_input = cat([control, mask_latents, masked_video_latents], dim=1), batch-of-2
        self.y = two(torch.cat([control_video_latents, mask_latents, masked_video_latents], dim=1))
        self.add = two(additional_control_latents)
        self.full_ref = two(ref_image_latents)
        self.context: List[torch.Tensor] = [u.to(dev, bf16) for u in list(negative_prompt_embeds) + list(prompt_embeds)]
        self.density = torch.full((2,), float(density), dtype=f32, device=dev)
        self.x_in = torch.empty((2,) + tuple(latents.shape[1:]), dtype=bf16, device=dev)
        self.launches = 0

    def step(self, i: int, t: float, sigma: float, sigma_next: float) -> None:
        self.tf.current_steps = i                                            # :847 (cfg_skip reads it)
        self.x_in[0].copy_(self.lat[0])                                      # :852 torch.cat([latents] * 2), bf16
        self.x_in[1].copy_(self.lat[0])
        ts = (self.tok_mask * float(t)).float().unsqueeze(0).expand(2, -1)   # :892-899 (seq_len == grid tokens)
        pred = self.tf(x=self.x_in, context=self.context, t=ts, density=self.density, seq_len=self.seq_len,
                       y=self.y, full_ref=self.full_ref, additional_control=self.add)
        # :926-934 in one launch: v = vu + s (vc - vu); lat += (sigma' - sigma) v; lat = (1-m) pinned + m lat
        ops.cfg_euler_step(pred[0], pred[1], self.guidance, float(sigma_next) - float(sigma), self.lat,
                           self.mask_full, self.pinned if self.repin else None)
        self.launches += 1

    def run(self, timesteps: Sequence[float], sigmas: Sequence[float],
            callback: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
        if len(sigmas) != len(timesteps) + 1:
            raise FlexamNativeError("DenoiseLoop.run: need one more sigma than timesteps (terminal sigma)")
        self.tf.num_inference_steps = len(timesteps)                         # :845
        for i, t in enumerate(timesteps):
            self.step(i, float(t), float(sigmas[i]), float(sigmas[i + 1]))
            if callback is not None:
                callback(i, self.lat)
        return self.lat.to(bf16)
