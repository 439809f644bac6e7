"""CPU tests of the host side of the native path: NativeEngine's launch sequence, buffer views and index tables are
run with the kernels replaced by their torch specifications (tests/cpu_ops_emul.py) and compared with the oracle.
The kernels themselves are checked on the GPU (tests/test_native_gpu.py)."""
import numpy as np
import pytest
import torch

import cpu_ops_emul
from flexam_b200.model import Wan2_2Transformer3DModel_FlexAM
from oracle import flexam_oracle as O
from oracle import synth


def build(cfg, device="cpu"):
    m = Wan2_2Transformer3DModel_FlexAM(
        model_type="ti2v", patch_size=cfg["patch_size"], text_len=cfg["text_len"], in_dim=cfg["in_dim"], dim=cfg["dim"],
        ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"], out_dim=cfg["out_dim"],
        num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], eps=cfg["eps"], add_ref_conv=True,
        in_dim_ref_conv=cfg["out_dim"], add_cnn_block=True, in_dim_cnn_block=cfg["in_dim_cnn"],
        out_dim_cnn_block=cfg["out_dim_cnn"], device=device)
    np_sd = synth.state_dict(cfg)
    m.load_state_dict({k: torch.from_numpy(v).to(torch.bfloat16) for k, v in np_sd.items()}, strict=True)
    return m, np_sd


def call(m, inp):
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    out = m(x=tt["x"].bfloat16(), t=tt["t"], context=[c.bfloat16() for c in ctx], seq_len=inp["seq_len"],
            y=tt["y"].bfloat16(), full_ref=tt["full_ref"].bfloat16(),
            additional_control=tt["additional_control"].bfloat16(), density=tt["density"])
    return out, tt, ctx


def rel(a, b):
    return (torch.linalg.vector_norm(a.float() - b.float()) / torch.linalg.vector_norm(b.float())).item()


@pytest.mark.parametrize("per_tok", [True, False])
def test_engine_host_logic_matches_oracle(monkeypatch, per_tok):
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, np_sd = build(cfg)
    inp = synth.inputs(cfg, 3, 8, 12, per_token_t=per_tok)
    out, tt, ctx = call(m, inp)
    assert out.shape == (2, 48, 3, 8, 12) and out.dtype == torch.bfloat16
    want = O.forward(O.to_torch_sd(np_sd), cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                     tt["additional_control"], tt["density"], policy="bf16")
    assert rel(out, want) < 4e-3
    assert m.engine().launches > 0


@pytest.mark.parametrize("table_cap", [64, 128])
def test_many_distinct_timesteps(monkeypatch, golden_dir, table_cap):
    """fg/bg-edit regime (fractional trilinear latent mask, pipeline :686-690, :891-898): 66 distinct per-token
    timesteps on this grid. Above the table capacity the engine runs the time MLP per token on the tensor cores and
    LayerNorm / gates read the per-token e0 (:444-446); below it the modulation tables hold one row per distinct value.
    Both must reproduce the REAL reference's output (tests/golden/tiny_frac.npz)."""
    import os
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    g = np.load(os.path.join(golden_dir, "tiny_frac.npz"))
    m, np_sd = build(cfg)
    eng = m.engine()
    if table_cap > 64:      # the dedup kernel lists at most 64 values; the emulation has no such limit
        eng.max_table_timesteps = table_cap
    inp = synth.inputs(cfg, 3, 8, 12, per_token_t="frac")
    assert len(np.unique(inp["t"])) == 66
    out, tt, ctx = call(m, inp)
    assert rel(out, torch.from_numpy(g["out"])) < 1e-2
    want = O.forward(O.to_torch_sd(np_sd), cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                     tt["additional_control"], tt["density"], policy="bf16")
    assert rel(out, want) < 4e-3
    assert eng.host_reads == 1


def test_cfg_skip_wrapper_halves_and_duplicates(monkeypatch):
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, _ = build(cfg)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True)
    full, _, _ = call(m, inp)
    m.enable_cfg_skip(0.5, 10)
    m.current_steps = 9                     # inside the last 50 % of the steps -> cond half only
    skipped, _, _ = call(m, inp)
    assert skipped.shape == full.shape
    assert torch.equal(skipped[0], skipped[1])
    assert rel(skipped[1], full[1]) < 3e-3   # CPU matmul blocking differs with the batch size
    m.current_steps = 0
    again, _, _ = call(m, inp)
    assert torch.equal(again, full)


def test_static_cache_tracks_in_place_edits(monkeypatch):
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, _ = build(cfg)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True)
    a, _, _ = call(m, inp)
    # LoRA-style in-place weight edit must invalidate the packed q|k|v copy
    with torch.no_grad():
        m.blocks[0].self_attn.q.weight.mul_(0.5)
    b, _, _ = call(m, inp)
    assert rel(b, a) > 1e-4


@pytest.mark.parametrize("per_tok,k_chunk", [(True, None), (False, 128)])
def test_precise_engine_host_logic_matches_fp32_oracle(monkeypatch, per_tok, k_chunk):
    """fp32 verification engine (flexam_b200/precise.py): plane splitting, plane-wise gathers and the launch order,
    with the kernels replaced by their torch specifications, against the fp32 oracle and the reference golden."""
    import os
    from flexam_b200.precise import precise_engine
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, np_sd = build(cfg)
    inp = synth.inputs(cfg, 3, 8, 12, per_token_t=per_tok)
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    eng = precise_engine(m)
    eng.k_chunk = k_chunk
    out = eng.forward(tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"], tt["additional_control"],
                      tt["density"])
    assert out.shape == (2, 48, 3, 8, 12) and out.dtype == torch.float32
    want = O.forward(O.to_torch_sd(np_sd), cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                     tt["additional_control"], tt["density"], policy="fp32")
    assert rel(out, want) < 2e-5
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "tiny_tok.npz" if per_tok else "tiny_sample.npz"))
    assert rel(out, torch.from_numpy(gold["out"])) < 1e-4


def test_sampling_loop_host_logic_matches_reference_golden(monkeypatch, golden_dir):
    """DenoiseLoop (CFG combine + Euler + re-pin kernel, TeaCache skipping, cfg_skip batch halving, static cache across
    steps) with emulated kernels vs the fixture made with the real reference module."""
    import loop_case
    cpu_ops_emul.install(monkeypatch)
    m, _ = build(synth.CONFIGS["tiny"])
    g = loop_case.golden(golden_dir)
    out, decisions, loop = loop_case.run_native_loop(m, g, "cpu")
    assert decisions == [bool(d) for d in g["decisions"]]
    assert out.dtype == torch.bfloat16 and loop.launches == 6
    # device->host reads of the whole 6-step loop: the TeaCache schedule (all decisions at once) and step 0's per-call
    # checks; steps 1..5 enqueue without host synchronisation (the reference loop syncs >= 3 times per step)
    assert loop.host_reads == 2
    r = rel(out, torch.from_numpy(g["out"]))
    print(f"emulated native loop vs reference loop golden: rel-L2 {r:.3e}")
    assert r < 1e-2
    # two bf16 paths with different rounding points, 6 guided steps (guidance 6 amplifies the CFG difference)
    want, _ = loop_case.run_oracle_loop(g, policy="bf16", dtype=torch.bfloat16)
    assert rel(out, want) < 1e-2


def test_install_honours_backend_switch(monkeypatch):
    """FLEXAM_BACKEND: 'reference' leaves the module's own forward in place, unknown values raise, and
    'native' rebinds forward (here with the kernels emulated) and reproduces the mirror class's result."""
    import flexam_b200.model as fx
    from flexam_b200.lib import FlexamNativeError
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, _ = build(cfg)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True)
    want, _, _ = call(m, inp)

    class Holder(torch.nn.Module):          # stands in for a reference module instance: same parameters and config
        def __init__(self, src):
            super().__init__()
            self.config = src.config
            self.cfg_skip_ratio, self.current_steps, self.num_inference_steps = None, 0, 1
            for name, p in src.named_parameters():
                fx._set_param(self, name, torch.nn.Parameter(p.detach().clone(), requires_grad=False))

        def forward(self, *a, **k):
            return "torch forward"

    monkeypatch.setenv("FLEXAM_BACKEND", "reference")
    h = fx.install(Holder(m))
    assert h.forward() == "torch forward" and not hasattr(h, "_flexam_engine")
    monkeypatch.setenv("FLEXAM_BACKEND", "triton")
    with pytest.raises(FlexamNativeError):
        fx.install(Holder(m))
    monkeypatch.setenv("FLEXAM_BACKEND", "native")
    h = fx.install(Holder(m))
    got, _, _ = call(h, inp)
    assert torch.equal(got, want)


def test_static_cache_is_keyed_by_content(monkeypatch):
    """The step-invariant cache (control fuser, context, cross K/V) must hit for a NEW tensor with the same values
    (the sampler builds its control batch with torch.cat every step) and miss when the values change, whatever the
    addresses: a second clip through the same engine equals a fresh engine's result."""
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, _ = build(cfg)
    a = synth.inputs(cfg, 2, 4, 8, per_token_t=True, tag="clipA")
    b = synth.inputs(cfg, 2, 4, 8, per_token_t=True, tag="clipB")
    eng = m.engine()
    calls = []
    real = eng._cnn_fuser
    monkeypatch.setattr(eng, "_cnn_fuser", lambda *x, **k: (calls.append(1), real(*x, **k))[1])
    out_a, _, _ = call(m, a)
    n = len(calls)
    out_a2, _, _ = call(m, {k: (v.copy() if hasattr(v, "copy") else v) for k, v in a.items()})   # same values, new tensors
    assert len(calls) == n and torch.equal(out_a, out_a2)
    out_b, _, _ = call(m, b)                                                    # other clip: must recompute
    assert len(calls) == 2 * n
    fresh, _ = build(cfg)
    want_b, _, _ = call(fresh, b)
    assert torch.equal(out_b, want_b) and not torch.equal(out_a, out_b)


def test_static_cache_identity_fast_path(monkeypatch):
    """Passing the very same tensors again must hit the cache without comparing contents (no device read-back); an
    in-place edit of one of them bumps its version and must miss."""
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, _ = build(cfg)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True)
    kw = dict(x=torch.from_numpy(inp["x"]).bfloat16(), t=torch.from_numpy(inp["t"]),
              context=[torch.from_numpy(c).bfloat16() for c in inp["context"]], seq_len=inp["seq_len"],
              y=torch.from_numpy(inp["y"]).bfloat16(), full_ref=torch.from_numpy(inp["full_ref"]).bfloat16(),
              additional_control=torch.from_numpy(inp["additional_control"]).bfloat16(),
              density=torch.from_numpy(inp["density"]))
    eng = m.engine()
    fuser_calls, probes = [], []
    real_fuser, real_probe = eng._cnn_fuser, eng._static_probe
    monkeypatch.setattr(eng, "_cnn_fuser", lambda *a, **k: (fuser_calls.append(1), real_fuser(*a, **k))[1])
    out1 = m(**kw)
    n = len(fuser_calls)
    monkeypatch.setattr(eng, "_static_probe", lambda *a, **k: (probes.append(real_probe(*a, **k)), probes[-1])[1])
    out2 = m(**kw)
    assert len(fuser_calls) == n and probes == ["hit"] and torch.equal(out1, out2)
    assert eng.host_reads == 1          # the single read of the call: weight fingerprint + number of distinct timesteps
    kw["additional_control"].mul_(0.5)                       # in place: same address, new version
    out3 = m(**kw)
    assert len(fuser_calls) == 2 * n and not torch.equal(out1, out3)


def test_from_pretrained_follows_the_reference_contract(tmp_path):
    """config.json + safetensors shards under path/subfolder, dict_mapping, patch-embedding channel padding, skipping of
    mis-sized tensors, dtype cast and the missing-config error (reference :1190-1332)."""
    import json
    from safetensors.torch import save_file
    cfg = synth.CONFIGS["tiny"]
    sd = {k: torch.from_numpy(v) for k, v in synth.state_dict(cfg).items()}
    root = tmp_path / "ckpt" / "transformer"
    root.mkdir(parents=True)
    conf = dict(_class_name="Wan2_2Transformer3DModel", _diffusers_version="0.33.0", model_type="ti2v",
                patch_size=[1, 2, 2], text_len=cfg["text_len"], in_dim=cfg["in_dim"] - 4, dim=cfg["dim"],
                ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"], out_dim=cfg["out_dim"],
                num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], eps=cfg["eps"], base_channels=cfg["in_dim"])
    (root / "config.json").write_text(json.dumps(conf))
    ckpt = dict(sd)
    ckpt["patch_embedding.weight"] = sd["patch_embedding.weight"][:, :cfg["in_dim"] - 4].contiguous()   # fewer channels
    ckpt["head.modulation"] = torch.zeros(1, 3, cfg["dim"])                                              # wrong size
    keys = sorted(ckpt)
    save_file({k: ckpt[k].contiguous() for k in keys[::2]}, str(root / "model-00001-of-00002.safetensors"))
    save_file({k: ckpt[k].contiguous() for k in keys[1::2]}, str(root / "model-00002-of-00002.safetensors"))
    extra = dict(add_ref_conv=True, in_dim_ref_conv=cfg["out_dim"], add_cnn_block=True, in_dim_cnn_block=cfg["in_dim_cnn"],
                 out_dim_cnn_block=cfg["out_dim_cnn"], dict_mapping={"base_channels": "in_dim"})
    m = Wan2_2Transformer3DModel_FlexAM.from_pretrained(str(tmp_path / "ckpt"), subfolder="transformer",
                                                        transformer_additional_kwargs=extra, torch_dtype=torch.bfloat16)
    got = m.state_dict()
    assert m.config.in_dim == cfg["in_dim"] and all(v.dtype == torch.bfloat16 for v in got.values())
    assert torch.equal(got["patch_embedding.weight"][:, :cfg["in_dim"] - 4], ckpt["patch_embedding.weight"].bfloat16())
    assert got["patch_embedding.weight"][:, cfg["in_dim"] - 4:].abs().max().item() == 0
    for k in ("blocks.1.ffn.2.weight", "cnn_conv3.0.bias", "time_projection.1.weight"):
        assert torch.equal(got[k], sd[k].bfloat16()), k
    assert "dict_mapping" in extra                                       # the caller's dict is not consumed
    # parameters the checkpoint does not carry get the reference's init_weights values, never torch.empty garbage: the
    # FlexAM additions of a base checkpoint (density MLPs, head.head.weight would be zero) — here head.modulation
    assert torch.isfinite(got["head.modulation"].float()).all() and got["head.modulation"].float().abs().max() < 1.0
    base = {k: v for k, v in ckpt.items() if not k.startswith("density_") and k != "head.modulation"}
    save_file({k: v.contiguous() for k, v in base.items()}, str(root / "model-00001-of-00002.safetensors"))
    (root / "model-00002-of-00002.safetensors").unlink()
    m2 = Wan2_2Transformer3DModel_FlexAM.from_pretrained(str(tmp_path / "ckpt"), subfolder="transformer",
                                                         transformer_additional_kwargs=extra)
    for k, v in m2.state_dict().items():
        if k.startswith("density_"):
            assert v.float().abs().max().item() == 0, k            # zero-initialised like the reference (:1172-1185)
    with pytest.raises(RuntimeError, match="config.json does not exist"):
        Wan2_2Transformer3DModel_FlexAM.from_pretrained(str(tmp_path / "ckpt"), subfolder="nope")


def test_weight_edits_are_picked_up(monkeypatch):
    """In-place torch ops on a parameter (version counter), replaced storages (`param.data = ...`, module.to()) and
    edits through `param.data` (the reference's LoRA merge, lora_utils.py:481-485; invisible to the counters, caught by
    the per-call weight fingerprint) all reach the packed q|k|v copy the kernels read and the cached cross-attention
    K/V; the result equals a freshly built model with the same weights."""
    import flexam_b200.model as fx
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, np_sd = build(cfg)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True)
    base, _, _ = call(m, inp)
    key = "blocks.1.self_attn.k.weight"
    p = dict(m.named_parameters())[key]
    delta = torch.from_numpy(synth.tensor("lora/delta", tuple(p.shape), 0.02)).bfloat16()

    def fresh(weight):
        f, _ = build(cfg)
        dict(f.named_parameters())[key].data.copy_(weight)
        return call(f, inp)[0]

    with torch.no_grad():
        p.add_(delta)                                           # version counter bumps: automatic
    want1 = fresh(p.data)
    got1, _, _ = call(m, inp)
    assert torch.equal(got1, want1) and not torch.equal(got1, base)
    p.data = (p.data.float() - delta.float()).bfloat16()       # new storage: automatic
    got2, _, _ = call(m, inp)
    assert torch.equal(got2, fresh(p.data))
    p.data += delta                                             # invisible to autograd: the fingerprint catches it
    got3, _, _ = call(m, inp)
    assert torch.equal(got3, fresh(p.data)) and not torch.equal(got3, got2)
    # a weight that feeds the CACHED cross-attention K/V (step-invariant cache): merge + unmerge through .data
    key2 = "blocks.0.cross_attn.v.weight"
    p2 = dict(m.named_parameters())[key2]
    d2 = torch.from_numpy(synth.tensor("lora/delta2", tuple(p2.shape), 0.05)).bfloat16()
    same_inputs = {k: torch.from_numpy(v) if hasattr(v, "dtype") else v for k, v in inp.items() if k != "context"}
    p2.data += d2
    got4, _, _ = call(m, inp)
    f, _ = build(cfg)
    dict(f.named_parameters())[key].data.copy_(p.data)
    dict(f.named_parameters())[key2].data.copy_(p2.data)
    assert torch.equal(got4, call(f, inp)[0]) and not torch.equal(got4, got3)
    fx.refresh(m)                                               # explicit form still works
    assert torch.equal(call(m, inp)[0], got4)


def test_engine_edge_shapes_on_one_engine(monkeypatch):
    """One engine, consecutive calls with different latent grids, batch sizes and prompt lengths — including a single
    latent frame, a batch of one, a batch of three, prompts of length 1 and of the full text_len — each against the
    oracle. Workspaces, RoPE position tables and the step-invariant caches are keyed by shape / content, so no call may
    see state left by another."""
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, np_sd = build(cfg)
    sd = O.to_torch_sd(np_sd)
    cases = [dict(F=3, H=8, W=12, B=2, prompt_lens=(37, 120)),
             dict(F=1, H=4, W=6, B=2, prompt_lens=(1, cfg["text_len"])),      # one latent frame (an image)
             dict(F=5, H=4, W=4, B=1, prompt_lens=(9,)),                      # no CFG batch
             dict(F=2, H=6, W=4, B=3, prompt_lens=(5, 6, 7)),                 # more than two samples
             dict(F=3, H=8, W=12, B=2, prompt_lens=(37, 120))]                # the first shape again, after the others
    outs = []
    for i, c in enumerate(cases):
        inp = synth.inputs(cfg, c["F"], c["H"], c["W"], B=c["B"], prompt_lens=c["prompt_lens"],
                           tag="in" if i in (0, 4) else f"edge{i}")
        out, tt, ctx = call(m, inp)
        want = O.forward(sd, cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                         tt["additional_control"], tt["density"], policy="bf16")
        assert out.shape == want.shape == (c["B"], cfg["out_dim"], c["F"], c["H"], c["W"]), (i, out.shape)
        assert torch.isfinite(out.float()).all() and rel(out, want) < 4e-3, (i, rel(out, want))
        outs.append(out)
    assert torch.equal(outs[0], outs[4])


@pytest.mark.parametrize("kind", ["fg_bg", "no_pin"])
def test_sampling_loop_with_other_masks(monkeypatch, kind):
    """DenoiseLoop beyond the full_edit fixture, 4 Euler steps against the restated pipeline loop (oracle/sampler_oracle.py)
    around the bf16-policy oracle forward: `fg_bg` — first frame pinned, the rest a fractional (trilinear-like) mask, so
    most tokens carry their own timestep (pipeline :686-690, :891-898) and the per-token time MLP path runs inside the
    loop; `no_pin` — the first latent frame is not fully masked out, so the loop must NOT re-pin it (:933)."""
    from flexam_b200.sampler import DenoiseLoop
    from oracle import make_golden, sampler_oracle
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    m, np_sd = build(cfg)
    sd = O.to_torch_sd(np_sd)
    F, H, W = 3, 8, 12
    lt = make_golden.loop_tensors(cfg, F, H, W)
    mask = torch.ones(1, 1, F, H, W)
    if kind == "fg_bg":
        ramp = torch.linspace(0, 1, W).view(1, 1, 1, 1, W) * torch.linspace(0.2, 1, H).view(1, 1, 1, H, 1)
        mask = (mask * ramp).bfloat16().float()
        mask[:, :, 0] = 0
    else:
        mask[:, :, 0, : H // 2] = 0                                      # half of the first frame stays editable
    lt["mask"] = mask
    lt["mask_latents"] = (1 - mask).expand(1, 4, F, H, W).contiguous()
    ts, sig = sampler_oracle.euler_schedule(4, 5.0)
    ts = synth.to_bf16_f32(ts)
    sig = np.concatenate([ts / np.float32(1000.0), np.zeros(1, np.float32)]).astype(np.float32)

    def oracle_loop(policy, dtype):
        def fn(x, context, t, density, seq_len, y, full_ref, additional_control):
            return O.forward(sd, cfg, x.float(), t.float(), [c.float() for c in context], seq_len, y.float(),
                             full_ref.float(), additional_control.float(), density.float(), policy=policy)
        return sampler_oracle.denoise_loop(fn, density=0.1, guidance_scale=6.0, timesteps=ts, sigmas=sig,
                                           set_step=lambda i, n: None, dtype=dtype, **lt)
    want, exact = oracle_loop("bf16", torch.bfloat16), oracle_loop("fp32", torch.float32)
    loop = DenoiseLoop(m, lt["latents"], lt["mask"], lt["masked_video_latents"], lt["mask_latents"],
                       lt["control_video_latents"], lt["additional_control"], lt["ref_image_latents"],
                       lt["negative_prompt_embeds"], lt["prompt_embeds"], density=0.1, guidance_scale=6.0)
    out = loop.run(ts, sig)
    assert loop.repin == (kind == "fg_bg")
    assert out.shape == want.shape and torch.isfinite(out.float()).all()
    # 4 guided steps of two different bf16 emulations (nothing pinned in `no_pin`: every token accumulates the noise):
    # the native loop must sit as close to the fp32 loop as the bf16-policy oracle loop does
    gap = rel(want, exact)
    assert rel(out, exact) < 1.25 * gap + 2e-3 and rel(out, want) < 2 * gap + 2e-3, (rel(out, exact), rel(out, want), gap)
    if kind == "fg_bg":
        assert torch.equal(out[:, :, 0].float(), lt["masked_video_latents"][:, :, 0].bfloat16().float())
