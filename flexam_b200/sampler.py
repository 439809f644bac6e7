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
    CNN control fuser, text embedding, cross K/V — hits on every step after the first).

    After the first step the loop enqueues WITHOUT host synchronisation: the per-token timesteps are ``mask * t`` with
    a constant mask, so their distinct values and inverse index are known up front (``t_dedup``); the TeaCache decisions
    depend on the timestep embedding only and are computed for all steps before the loop (one read-back,
    ``teacache_schedule``); the engine is told that weights and control inputs cannot change inside the loop
    (``trusted``). ``graph=True`` additionally captures the transformer call of each (batch size, run / skip) variant
    in a CUDA graph at its second occurrence and replays it afterwards (single-GPU engines only)."""

    def __init__(self, transformer, latents: torch.Tensor, mask: torch.Tensor, masked_video_latents: torch.Tensor,
                 mask_latents: torch.Tensor, control_video_latents: torch.Tensor,
                 additional_control_latents: torch.Tensor, ref_image_latents: torch.Tensor,
                 negative_prompt_embeds: Sequence[torch.Tensor], prompt_embeds: Sequence[torch.Tensor],
                 density: float, guidance_scale: float = 6.0, graph: bool = False):
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
        self.host_reads = 0          # device->host reads made by the transformer calls of the last run()
        # distinct mask factors and their inverse index: ONE read-back per loop instead of a torch.unique per step
        vals, inv = torch.unique(self.tok_mask.float(), return_inverse=True)
        self.mask_vals = vals.to(bf16)                                        # [U]
        self.mask_inv = inv.to(torch.int32).view(1, -1).expand(2, -1).contiguous()   # [2, L0]
        self.ts = torch.empty((2, self.seq_len), dtype=f32, device=dev)       # per-token timesteps of the current step
        self.uniq = torch.empty((self.mask_vals.numel(),), dtype=f32, device=dev)
        self.use_graph = bool(graph)
        self._graphs = {}            # (batch, run_blocks) -> [occurrences, CUDAGraph | None, static output]
        self.graph_replays = 0
        self.teacache_distances: List[float] = []
        self.decisions: List[bool] = []   # per step: did the block stack run (False = TeaCache re-applied the residual)

    def _engine(self):
        return self.tf.engine() if hasattr(self.tf, "engine") else self.tf._flexam_engine

    def teacache_schedule(self, timesteps: Sequence[float]) -> Optional[List[bool]]:
        """TeaCache decisions (:978-1000) of every step of the loop, before the loop: the decision input is the
        timestep embedding ``e0`` of the LAST token, a function of ``mask[-1] * t_i`` and the weights only. One batched
        pass of the fp32 time MLP over all steps and ONE read-back of the n-1 relative-L1 distances; the accumulate /
        rescale / threshold recurrence then runs on the host exactly as the reference does."""
        tc = getattr(self.tf, "teacache", None)
        if tc is None:
            return None
        eng = self._engine()
        n = len(timesteps)
        t_last = torch.stack([(self.tok_mask[-1:] * float(t)).float() for t in timesteps]).view(-1)     # bf16 products
        e0 = torch.cat([eng._embed_mlp("time_embedding", "time_projection", t_last[i:i + 16].contiguous())[1]
                        for i in range(0, n, 16)])                                       # [n, 6D], fp32 kernel path
        prev0 = tc.previous_modulated_input
        if prev0 is not None and tc.cnt >= tc.num_skip_start_steps:
            e0p = torch.cat([prev0.reshape(-1, e0.shape[1])[:1].to(e0), e0])
        else:
            e0p = torch.cat([e0[:1], e0])
        d = ((e0p[1:] - e0p[:-1]).abs().mean(dim=1) / e0p[:-1].abs().mean(dim=1)).tolist()          # the one read-back
        self.host_reads += 1
        self.teacache_distances = d            # relative L1 distance of consecutive steps' modulated inputs (diagnostics)
        acc, cnt, out = float(tc.accumulated_rel_l1_distance), int(tc.cnt), []
        for i in range(n):
            if cnt < tc.num_skip_start_steps:
                should, acc = True, 0.0
            else:
                acc += float(tc.rescale_func(d[i]))
                if acc < tc.rel_l1_thresh:
                    should = False
                else:
                    should, acc = True, 0.0
            out.append(should)
            cnt += 1
            if cnt == tc.num_steps:        # the forward resets the cache there (:1119-1122)
                acc, cnt = 0.0, 0
        self._tc_final = (acc, e0[-1:].clone())
        return out

    def _call(self, decision: Optional[bool]):
        eng = self._engine()
        kw = {"t_dedup": (self.uniq, self.mask_inv)}
        if decision is not None:
            kw["teacache_decision"] = decision
        self.tf._fx_loop_kwargs = kw
        try:
            return self.tf(x=self.x_in, context=self.context, t=self.ts, density=self.density, seq_len=self.seq_len,
                           y=self.y, full_ref=self.full_ref, additional_control=self.add)
        finally:
            self.tf._fx_loop_kwargs = {}
            self.host_reads += eng.host_reads

    def _skipping_cfg(self, i: int) -> bool:
        r = getattr(self.tf, "cfg_skip_ratio", None)
        n = getattr(self.tf, "num_inference_steps", None)
        return r is not None and n is not None and i >= n * (1 - r)

    def step(self, i: int, t: float, sigma: float, sigma_next: float, decision: Optional[bool] = None) -> None:
        self.tf.current_steps = i                                            # :847 (cfg_skip reads it)
        self.x_in[0].copy_(self.lat[0])                                      # :852 torch.cat([latents] * 2), bf16
        self.x_in[1].copy_(self.lat[0])
        self.uniq.copy_((self.mask_vals * float(t)).float())                 # :892-899 distinct values of mask * t
        self.ts.copy_((self.tok_mask * float(t)).float().unsqueeze(0).expand(2, -1))
        eng = self._engine()
        graphable = (self.use_graph and eng.par is None and eng.trusted and
                     (decision is not None or getattr(self.tf, "teacache", None) is None))
        if not graphable:
            pred = self._call(decision)
        else:
            key = (1 if self._skipping_cfg(i) else 2, decision)
            slot = self._graphs.setdefault(key, [0, None, None])
            slot[0] += 1
            if slot[0] == 1:                 # first occurrence: eager (allocates the engine's workspaces)
                pred = self._call(decision)
            else:
                tc = getattr(self.tf, "teacache", None)
                if slot[1] is None:          # second occurrence: capture (records, does not execute) ...
                    state = (tc.cnt, tc.should_calc) if tc is not None else None
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        slot[2] = self._call(decision)
                    slot[1] = g
                    if tc is not None:       # the captured call ran the host bookkeeping once; the replay below is the step
                        tc.cnt, tc.should_calc = state
                slot[1].replay()             # ... then replay
                self.graph_replays += 1
                if tc is not None:           # host bookkeeping the captured forward would have done (:1119-1122)
                    tc.should_calc = bool(decision)
                    tc.cnt += 1
                    if tc.cnt == tc.num_steps:
                        tc.reset()
                pred = slot[2]
        tc_obj = getattr(self.tf, "teacache", None)
        self.decisions.append(bool(decision) if decision is not None else
                              (bool(tc_obj.should_calc) if tc_obj is not None else True))
        # :926-934 in one launch: v = vu + s (vc - vu); lat += (sigma' - sigma) v; lat = (1-m) pinned + m lat
        ops.cfg_euler_step(pred[0], pred[1], self.guidance, float(sigma_next) - float(sigma), self.lat,
                           self.mask_full, self.pinned if self.repin else None)
        self.launches += 1

    def run(self, timesteps: Sequence[float], sigmas: Sequence[float],
            callback: Optional[Callable[[int, torch.Tensor], None]] = None) -> torch.Tensor:
        if len(sigmas) != len(timesteps) + 1:
            raise FlexamNativeError("DenoiseLoop.run: need one more sigma than timesteps (terminal sigma)")
        self.tf.num_inference_steps = len(timesteps)                         # :845
        self.host_reads = 0
        self.decisions = []
        eng = self._engine()
        decisions = self.teacache_schedule(timesteps)
        was_trusted = eng.trusted
        try:
            for i, t in enumerate(timesteps):
                if i == 1:   # step 0 ran every per-call check (weights, control inputs); they cannot change inside the loop
                    eng.trusted = True
                self.step(i, float(t), float(sigmas[i]), float(sigmas[i + 1]),
                          None if decisions is None else decisions[i])
                if callback is not None:
                    callback(i, self.lat)
        finally:
            eng.trusted = was_trusted
        tc = getattr(self.tf, "teacache", None)
        if tc is not None and decisions is not None and tc.cnt != 0:   # loop ended mid-schedule: leave a consistent state
            tc.accumulated_rel_l1_distance, tc.previous_modulated_input = self._tc_final
        return self.lat.to(bf16)
