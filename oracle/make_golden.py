"""TEST INFRASTRUCTURE — generates tests/golden/*.npz by running the REAL reference module (build container only).

    python oracle/make_golden.py            # writes tests/golden/{tiny_tok,tiny_sample,real2_tok,tiny_loop}.npz
    python oracle/make_golden.py real_tok   # the full-depth fixture (minutes, ~45 GB RAM)

Each fixture stores the reference's fp32 CPU output for ``oracle.synth`` weights/inputs (which are regenerated
from names, so they are not stored), plus a few small intermediate slices used to localise a mismatch.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import flexam_oracle as O  # noqa: E402
from oracle import ref_import, synth  # noqa: E402

CASES = {
    # name: (config, (F, H, W), per-token timesteps)
    "tiny_tok": ("tiny", (3, 8, 12), True),
    "tiny_sample": ("tiny", (3, 8, 12), False),
    "real2_tok": ("real2", (5, 16, 28), True),   # BASELINE config-1 grid (560 + 112 tokens), 2 of 30 layers
    # BASELINE config 1 itself: the full 30-layer, 5.0 B-parameter model at 17 frames 256x448 (fp32 on CPU: ~6 min of
    # weight hashing + 20 s forward, ~45 GB of RAM). Not part of the default list: `make_golden.py real_tok`.
    "real_tok": ("real", (5, 16, 28), True),
    # fg/bg-edit regime: fractional (trilinear) latent mask => (almost) every token has its own timestep
    "tiny_frac": ("tiny", (3, 8, 12), "frac"),
    "real2_frac": ("real2", (5, 16, 28), "frac"),
}
DEFAULT = ("tiny_tok", "tiny_sample", "real2_tok", "tiny_frac", "real2_frac", "tiny_loop", "rope_tables", "t5_tiny",
           "t5_real2", "vae_tiny")


def run_case(name: str):
    cfg_name, (F, H, W), per_tok = CASES[name]
    cfg = synth.CONFIGS[cfg_name]
    t0 = time.time()
    model = ref_import.build_reference_model(cfg).eval()
    sd = O.to_torch_sd(synth.state_dict(cfg))
    model.load_state_dict(sd, strict=True)
    inp = synth.inputs(cfg, F, H, W, per_token_t=per_tok)
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    with torch.no_grad():
        out = model(x=tt["x"], t=tt["t"], context=ctx, seq_len=inp["seq_len"], y=tt["y"], full_ref=tt["full_ref"],
                    additional_control=tt["additional_control"], density=tt["density"])
        taps = {}
        mine = O.forward(sd, cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                         tt["additional_control"], tt["density"], taps=taps)
    rel = ((out - mine).norm() / out.norm()).item()
    print(f"{name}: reference out {tuple(out.shape)} |max| {out.abs().max():.3f}  oracle rel-L2 {rel:.2e}  "
          f"({time.time() - t0:.1f}s)")
    assert rel < 2e-5, "oracle restatement disagrees with the reference"
    rows = slice(0, None, max(1, taps["x0"].shape[1] // 16))
    np.savez_compressed(
        os.path.join(ROOT, "tests", "golden", name + ".npz"),
        out=out.numpy().astype(np.float32),
        # oracle-side intermediates (validated end-to-end by the assert above), sample 0, every ~L/16-th token
        x0_rows=taps["x0"][0, rows].numpy(), after_self0_rows=taps["b0.after_self"][rows].numpy(),
        after_cross0_rows=taps["b0.after_cross"][rows].numpy(), x_final_rows=taps["x_final"][rows].numpy(),
        cnn_out_slice=taps["cnn_out"][0, :, 0].numpy(), ctx_rows=taps["ctx"][0, ::64].numpy(),
        meta=np.array([F, H, W, 2 if per_tok == "frac" else int(per_tok)], dtype=np.int64), config=np.array(cfg_name))


# the sampling-loop fixture: 6 Euler steps (shift 5, guidance 6, density 10 = `full_edit`) around the REAL reference
# module with TeaCache and cfg_skip enabled, so skipped block stacks and halved batches are both on the path
LOOP = dict(config="tiny", grid=(3, 8, 12), steps=6, shift=5.0, guidance=6.0, density=10.0, cfg_skip_ratio=0.34,
            # synthetic weights move the timestep embedding by a relative L1 of ~1 per step, so the rescale polynomial
            # is the identity and the threshold sits between one and two steps' worth (with an 8 % margin)
            teacache=dict(coefficients=[1.0, 0.0], rel_l1_thresh=1.9, num_skip_start_steps=1))


def loop_tensors(cfg, F, H, W):
    li = synth.loop_inputs(cfg, F, H, W)
    return {k: ([torch.from_numpy(u) for u in v] if isinstance(v, list) else torch.from_numpy(v)) for k, v in li.items()}


def run_loop(name: str = "tiny_loop"):
    from oracle import sampler_oracle as S
    cfg = synth.CONFIGS[LOOP["config"]]
    F, H, W = LOOP["grid"]
    sd = O.to_torch_sd(synth.state_dict(cfg))
    model = ref_import.build_reference_model(cfg).eval()
    model.load_state_dict(sd, strict=True)
    tc = LOOP["teacache"]
    model.enable_teacache(tc["coefficients"], LOOP["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"], offload=False)
    model.enable_cfg_skip(LOOP["cfg_skip_ratio"], LOOP["steps"])
    ts, sig = S.euler_schedule(LOOP["steps"], LOOP["shift"])
    # custom timesteps (the pipeline accepts them): the schedule rounded to bf16-representable values, so that the
    # bf16 pipeline's `mask * t` (which rounds t to bf16) and this fp32 run see identical timesteps
    ts = synth.to_bf16_f32(ts)
    sig = np.concatenate([ts / np.float32(1000.0), np.zeros(1, np.float32)]).astype(np.float32)
    decisions, batches = [], []

    def ref_fn(**kw):
        batches.append(len(kw["x"]))
        out = model(**kw)
        decisions.append(bool(model.should_calc))
        return out

    def set_ref(i, n):
        model.current_steps, model.num_inference_steps = i, n

    lt = loop_tensors(cfg, F, H, W)
    trace = []
    with torch.no_grad():
        ref = S.denoise_loop(ref_fn, density=LOOP["density"], guidance_scale=LOOP["guidance"], timesteps=ts, sigmas=sig,
                             set_step=set_ref, trace=trace, **lt)
        # the same loop around the oracle forward (TeaCache / cfg_skip restated) must agree with it
        otc = O.TeaCacheOracle(tc["coefficients"], LOOP["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"])
        state = {}

        def set_or(i, n):
            state["i"], state["n"] = i, n

        def or_fn(x, context, t, density, seq_len, y, full_ref, additional_control):
            return O.forward_cfg_skip(sd, cfg, x, t, context, seq_len, y, full_ref, additional_control, density,
                                      cfg_skip_ratio=LOOP["cfg_skip_ratio"], current_step=state["i"],
                                      num_steps=state["n"], teacache=otc)
        mine = S.denoise_loop(or_fn, density=LOOP["density"], guidance_scale=LOOP["guidance"], timesteps=ts, sigmas=sig,
                              set_step=set_or, **lt)
    rel = ((ref - mine).norm() / ref.norm()).item()
    print(f"{name}: TeaCache decisions {decisions} (oracle {otc.decisions}), forward batch sizes seen by cfg_skip "
          f"wrapper input {batches}; oracle-loop rel-L2 {rel:.2e}")
    assert decisions == otc.decisions and not all(decisions) and any(decisions[1:]), "want a mixed skip pattern"
    assert rel < 2e-5
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), out=ref.numpy().astype(np.float32),
                        decisions=np.array(decisions), timesteps=ts, sigmas=sig,
                        step_norms=np.array([t.norm().item() for t in trace], dtype=np.float64),
                        step0=trace[0].numpy().astype(np.float32))


T5_CASES = {
    # name: (t5 config, padded length, prompt lengths)
    "t5_tiny": ("tiny", 64, (11, 40)),
    "t5_real2": ("real2", 512, (37, 120)),      # real width and head layout, 2 of 24 layers, the pipeline's 512-token padding
}


def run_t5(name: str):
    """The REAL umT5 encoder (FlexAM/models/wan_text_encoder.py:256-304) on synthetic weights and token ids, fp32 CPU."""
    from oracle import t5_oracle as T
    cfg_name, L, lens = T5_CASES[name]
    cfg = T.T5_CONFIGS[cfg_name]
    t0 = time.time()
    model = ref_import.build_reference_t5(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in T.state_dict(cfg).items()}
    model.load_state_dict(sd, strict=True)
    ids, mask = T.inputs(cfg, L=L, lens=lens)
    with torch.no_grad():
        ref = model(torch.from_numpy(ids), torch.from_numpy(mask))[0]
        mine = T.forward(sd, cfg, torch.from_numpy(ids), torch.from_numpy(mask))
    rel = ((ref - mine).norm() / ref.norm()).item()
    print(f"{name}: reference out {tuple(ref.shape)} |max| {ref.abs().max():.3f}  oracle rel-L2 {rel:.2e}  ({time.time() - t0:.1f}s)")
    assert rel < 2e-5, "T5 oracle restatement disagrees with the reference"
    step = 4 if L >= 256 else 1          # every 4th token row of the long fixture keeps the file small
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                        out=ref[:, ::step].numpy().astype(np.float32),
                        meta=np.array([L, step] + list(lens), dtype=np.int64), config=np.array(cfg_name))


VAE_CASES = {
    # name: (vae config, latent grid (T, H, W)); 3 latent frames cover the first chunk, the "Rep" rule and the cached path
    "vae_tiny": ("tiny", (3, 4, 6)),
}


def run_vae(name: str):
    """The REAL Wan2.2 VAE decoder (FlexAM/models/wan_vae3_8.py:820-849 + the wrapper's clamp :1043) on synthetic weights."""
    from oracle import vae_oracle as V
    cfg_name, (T, H, W) = VAE_CASES[name]
    cfg = V.VAE_CONFIGS[cfg_name]
    t0 = time.time()
    model = ref_import.build_reference_vae(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in {**V.encoder_state_dict(cfg), **V.state_dict(cfg)}.items()}
    model.load_state_dict(sd, strict=True)
    z, scale = torch.from_numpy(V.latents(cfg, T, H, W)), V.latent_scale(cfg)
    x = torch.from_numpy(V.video(cfg, 1 + 4 * (T - 1), 16 * H, 16 * W))
    with torch.no_grad():
        ref = model.decode(z, scale).clamp_(-1, 1)
        mine = V.decode(sd, cfg, z, scale)
        ref_e = model.encode(x, scale)
        mine_e = V.encode(sd, cfg, x, scale)
    rel, rel_e = ((ref - mine).norm() / ref.norm()).item(), ((ref_e - mine_e).norm() / ref_e.norm()).item()
    print(f"{name}: reference decode {tuple(ref.shape)} oracle rel-L2 {rel:.2e}; encode {tuple(ref_e.shape)} oracle rel-L2 "
          f"{rel_e:.2e}  ({time.time() - t0:.1f}s)")
    assert rel < 2e-5 and rel_e < 2e-5, "VAE oracle disagrees with the reference"
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), out=ref.numpy().astype(np.float32),
                        enc=ref_e.numpy().astype(np.float32), meta=np.array([T, H, W], dtype=np.int64),
                        config=np.array(cfg_name))


def run_rope(name: str = "rope_tables"):
    """RoPE tables of the REAL reference module, default and after enable_riflex() with its default arguments
    (:774-788): a few position rows of the complex128 / complex64 tables as float64 (cos, sin)."""
    cfg = synth.CONFIGS["tiny"]
    model = ref_import.build_reference_model(cfg)
    rows = np.array([0, 1, 2, 7, 24, 65, 66, 511, 1023])
    plain = torch.view_as_real(model.freqs.to(torch.complex128))[rows].numpy()
    model.enable_riflex()                      # k = 6, L_test = 66, L_test_scale = 4.886
    riflex = torch.view_as_real(model.freqs.to(torch.complex128))[rows].numpy()
    model.enable_riflex(k=4, L_test=49, L_test_scale=None)
    riflex2 = torch.view_as_real(model.freqs.to(torch.complex128))[rows].numpy()
    assert not np.allclose(plain, riflex) and plain.shape == (len(rows), 64, 2)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"), rows=rows, plain=plain, riflex=riflex,
                        riflex_k4_L49=riflex2)
    print(f"{name}: head_dim {model.d if hasattr(model, 'd') else 128}, {len(rows)} rows, "
          f"max |plain - riflex| {np.abs(plain - riflex).max():.3f}")


if __name__ == "__main__":
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    for n in (sys.argv[1:] or DEFAULT):
        (run_loop(n) if n == "tiny_loop" else run_rope(n) if n == "rope_tables" else run_t5(n) if n in T5_CASES
         else run_vae(n) if n in VAE_CASES else run_case(n))
