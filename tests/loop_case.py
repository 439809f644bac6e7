"""TEST INFRASTRUCTURE — the sampling-loop fixture (tests/golden/tiny_loop.npz, made by oracle/make_golden.py with the
REAL reference module inside the restated pipeline loop): 6 Euler steps with TeaCache and cfg_skip enabled."""
import os

import numpy as np
import torch

from oracle import flexam_oracle as O
from oracle import make_golden, sampler_oracle, synth

LOOP = make_golden.LOOP


def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "tiny_loop.npz"))


def run_oracle_loop(g, policy="fp32", dtype=torch.float32, device="cpu"):
    cfg = synth.CONFIGS[LOOP["config"]]
    sd = O.to_torch_sd(synth.state_dict(cfg), device)
    tc = LOOP["teacache"]
    otc = O.TeaCacheOracle(tc["coefficients"], LOOP["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"])
    state = {}

    def fn(x, context, t, density, seq_len, y, full_ref, additional_control):
        return O.forward_cfg_skip(sd, cfg, x.float(), t.float(), [c.float() for c in context], seq_len, y.float(),
                                  full_ref.float(), additional_control.float(), density.float(),
                                  cfg_skip_ratio=LOOP["cfg_skip_ratio"], current_step=state["i"], num_steps=state["n"],
                                  teacache=otc, policy=policy)
    lt = make_golden.loop_tensors(cfg, *LOOP["grid"])
    lt = {k: ([u.to(device) for u in v] if isinstance(v, list) else v.to(device)) for k, v in lt.items()}
    out = sampler_oracle.denoise_loop(fn, density=LOOP["density"], guidance_scale=LOOP["guidance"],
                                      timesteps=g["timesteps"], sigmas=g["sigmas"],
                                      set_step=lambda i, n: state.update(i=i, n=n), dtype=dtype, **lt)
    return out, otc.decisions


def run_native_loop(model, g, device, graph=False):
    """flexam_b200.sampler.DenoiseLoop around the (native or kernel-emulated) mirror module."""
    from flexam_b200.sampler import DenoiseLoop
    cfg = synth.CONFIGS[LOOP["config"]]
    tc = LOOP["teacache"]
    model.enable_teacache(tc["coefficients"], LOOP["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"], offload=False)
    model.enable_cfg_skip(LOOP["cfg_skip_ratio"], LOOP["steps"])
    lt = make_golden.loop_tensors(cfg, *LOOP["grid"])
    lt = {k: ([u.to(device) for u in v] if isinstance(v, list) else v.to(device)) for k, v in lt.items()}
    loop = DenoiseLoop(model, lt["latents"], lt["mask"], lt["masked_video_latents"], lt["mask_latents"],
                       lt["control_video_latents"], lt["additional_control"], lt["ref_image_latents"],
                       lt["negative_prompt_embeds"], lt["prompt_embeds"], density=LOOP["density"],
                       guidance_scale=LOOP["guidance"], graph=graph)
    out = loop.run(g["timesteps"], g["sigmas"])
    return out, list(loop.decisions), loop
