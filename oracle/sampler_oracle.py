"""TEST INFRASTRUCTURE — CPU restatement of the sampling loop around the transformer
(FlexAM/pipeline/pipeline_wan2_2_fun_control_FlexAM.py:843-934) for the FlexAM 5B configuration, with the model
call abstracted as ``model_fn`` so the SAME loop can drive the real reference module (``oracle/make_golden.py``,
build container only), the oracle forward, or be compared with ``flexam_b200.sampler.DenoiseLoop``.

Parity status: the loop body is restated from the pipeline lines cited below and pinned end-to-end by
``tests/golden/tiny_loop.npz`` (real reference module inside this loop). The scheduler arithmetic
(diffusers ``FlowMatchEulerDiscreteScheduler``, absent from the reference tree and from this image) is restated
from its published semantics — "parity unpinned" for that class alone. Only tests/, smoke() and bench.py's CPU
baseline may import this file.
"""
from __future__ import annotations

from typing import Callable, List, Sequence

import numpy as np
import torch


def euler_schedule(num_inference_steps: int, shift: float = 5.0, num_train_timesteps: int = 1000):
    """FlowMatchEulerDiscreteScheduler(shift, use_dynamic_shifting=False): __init__ then set_timesteps(n)."""
    N = num_train_timesteps
    sig = (np.linspace(1, N, N, dtype=np.float32)[::-1].copy() / N).astype(np.float32)
    sig = shift * sig / (1 + (shift - 1) * sig)
    t_max, t_min = float(sig[0]) * N, float(sig[-1]) * N                  # _sigma_to_t(sigma_max / sigma_min)
    s = np.linspace(t_max, t_min, num_inference_steps) / N
    s = (shift * s / (1 + (shift - 1) * s)).astype(np.float32)
    return (s * N).astype(np.float32), np.concatenate([s, np.zeros(1, dtype=np.float32)])


def denoise_loop(model_fn: Callable, latents: torch.Tensor, mask: torch.Tensor, masked_video_latents: torch.Tensor,
                 mask_latents: torch.Tensor, control_video_latents: torch.Tensor, additional_control: torch.Tensor,
                 ref_image_latents: torch.Tensor, negative_prompt_embeds: List[torch.Tensor],
                 prompt_embeds: List[torch.Tensor], density: float, guidance_scale: float, timesteps: Sequence[float],
                 sigmas: Sequence[float], set_step: Callable[[int, int], None], dtype=torch.float32,
                 trace: list = None) -> torch.Tensor:
    """Returns the final latents [1, C, F, H, W] in `dtype` (the pipeline's weight_dtype)."""
    wd = dtype
    latents = latents.to(wd)
    mask = mask.to(wd)
    masked = masked_video_latents.to(wd)
    pin = not bool(mask[:, :, 0, :, :].any())
    if pin:                                                               # :688-690
        latents = (1 - mask) * masked + mask * latents
    _, _, F, H, W = latents.shape
    seq_len = F * (H // 2) * (W // 2)                                     # :838-839
    context = list(negative_prompt_embeds) + list(prompt_embeds)         # :598-599
    n = len(timesteps)
    for i in range(n):
        t = torch.tensor(float(timesteps[i]), dtype=torch.float32)
        set_step(i, n)                                                    # :845, :847
        x_in = torch.cat([latents] * 2)                                   # :852
        ctrl = torch.cat([control_video_latents] * 2).to(wd)              # :865-867
        add = torch.cat([additional_control] * 2).to(wd)                  # :869-870
        y = torch.cat([torch.cat([mask_latents] * 2), torch.cat([masked] * 2)], dim=1).to(wd)   # :872-877
        ctrl = torch.cat([ctrl, y], dim=1)                                # :878-879
        full_ref = torch.cat([ref_image_latents] * 2).to(wd)              # :887-890
        temp_ts = (mask[0][0][:, ::2, ::2] * t).flatten()                 # :892
        temp_ts = torch.cat([temp_ts, temp_ts.new_ones(seq_len - temp_ts.size(0)) * t]).unsqueeze(0)
        timestep = temp_ts.expand(2, temp_ts.size(1))                     # :897-898
        dens = torch.tensor([float(density)]).expand(2)                   # :901
        pred = model_fn(x=x_in, context=context, t=timestep, density=dens, seq_len=seq_len, y=ctrl,
                        full_ref=full_ref, additional_control=add).to(wd)
        pu, pc = pred.chunk(2)                                            # :927-928
        pred = pu + guidance_scale * (pc - pu)
        # scheduler.step :931 — fp32 update, result cast to the model output dtype
        # (dt is a 0-dim fp32 tensor: under torch's promotion rules dt * pred stays in pred's dtype)
        dt = torch.tensor(float(sigmas[i + 1]), dtype=torch.float32) - torch.tensor(float(sigmas[i]), dtype=torch.float32)
        prev = latents.to(torch.float32) + dt * pred
        latents = prev.to(pred.dtype)
        if pin:                                                           # :933-934
            latents = (1 - mask) * masked + mask * latents
        if trace is not None:
            trace.append(latents.float().clone())
    return latents
