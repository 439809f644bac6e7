"""CPU checks that pin the oracle: against the committed golden outputs of the REAL reference module, and against
the live reference when /root/reference is present (build container only)."""
import os

import numpy as np
import pytest
import torch

from oracle import flexam_oracle as O
from oracle import ref_import, synth


def _run_oracle(cfg_name, grid, per_tok, policy="fp32", taps=None):
    cfg = synth.CONFIGS[cfg_name]
    sd = O.to_torch_sd(synth.state_dict(cfg))
    inp = synth.inputs(cfg, *grid, per_token_t=per_tok)
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    with torch.no_grad():
        return O.forward(sd, cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                         tt["additional_control"], tt["density"], policy=policy, taps=taps)


def _rel(a, b):
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b)).item()


def _tmode(meta_flag):
    """meta[3] of a fixture: 0 per-sample timesteps, 1 per-token (two values), 2 per-token fractional (fg/bg edit)."""
    return "frac" if int(meta_flag) == 2 else bool(meta_flag)


@pytest.mark.parametrize("name", ["tiny_tok", "tiny_sample", "tiny_frac"])
def test_oracle_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    F, H, W, per_tok = (int(v) for v in g["meta"])
    per_tok = _tmode(per_tok)
    taps = {}
    out = _run_oracle(str(g["config"]), (F, H, W), per_tok, taps=taps)
    assert _rel(out, torch.from_numpy(g["out"])) < 2e-5
    rows = slice(0, None, max(1, taps["x0"].shape[1] // 16))
    assert _rel(taps["x_final"][rows], torch.from_numpy(g["x_final_rows"])) < 2e-5


@pytest.mark.slow
def test_oracle_matches_reference_golden_real_width(golden_dir):
    g = np.load(os.path.join(golden_dir, "real2_tok.npz"))
    F, H, W, per_tok = (int(v) for v in g["meta"])
    out = _run_oracle("real2", (F, H, W), bool(per_tok))
    assert _rel(out, torch.from_numpy(g["out"])) < 2e-5


def test_bf16_policy_stays_within_the_bf16_gate():
    """The emulated autocast flow must sit inside the north-star tolerance (rel-L2 <= 1e-2 vs fp32)."""
    ref = _run_oracle("tiny", (3, 8, 12), True)
    emu = _run_oracle("tiny", (3, 8, 12), True, policy="bf16")
    assert 1e-5 < _rel(emu, ref) < 1e-2


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_oracle_matches_live_reference():
    cfg = synth.CONFIGS["tiny"]
    model = ref_import.build_reference_model(cfg).eval()
    sd = O.to_torch_sd(synth.state_dict(cfg))
    model.load_state_dict(sd, strict=True)
    inp = synth.inputs(cfg, 2, 4, 8, per_token_t=True, tag="live")
    ctx = [torch.from_numpy(c) for c in inp["context"]]
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    with torch.no_grad():
        ref = model(x=tt["x"], t=tt["t"], context=ctx, seq_len=inp["seq_len"], y=tt["y"], full_ref=tt["full_ref"],
                    additional_control=tt["additional_control"], density=tt["density"])
        mine = O.forward(sd, cfg, tt["x"], tt["t"], ctx, inp["seq_len"], tt["y"], tt["full_ref"],
                         tt["additional_control"], tt["density"])
    assert _rel(mine, ref) < 2e-5


def test_rope_table_matches_reference_construction():
    a = O.rope_angles(128)
    assert a.shape == (1024, 64) and a.dtype == torch.float64
    # frame axis uses 44-wide frequencies, row/col 42-wide (d - 4*(d//6), 2*(d//6))
    assert torch.allclose(a[1, :22], 1.0 / torch.pow(10000.0, torch.arange(0, 44, 2, dtype=torch.float64) / 44))
    assert torch.allclose(a[1, 22:43], 1.0 / torch.pow(10000.0, torch.arange(0, 42, 2, dtype=torch.float64) / 42))
    from flexam_b200.model import rope_table
    assert torch.equal(rope_table(128), O.rope_table_f32(128))


def test_param_tree_matches_between_product_and_oracle():
    from flexam_b200.model import param_shapes
    cfg = dict(synth.CONFIGS["tiny"])
    prod = param_shapes(dict(cfg, in_dim_ref_conv=cfg["out_dim"], in_dim_cnn_block=cfg["in_dim_cnn"],
                             out_dim_cnn_block=cfg["out_dim_cnn"]))
    orc = {n: tuple(s) for n, s, _, _ in synth.param_specs(cfg)}
    assert prod == orc


def test_oracle_sampling_loop_matches_reference_golden(golden_dir):
    """The restated pipeline loop + TeaCache + cfg_skip around the oracle forward reproduces the fixture made with the
    real reference module (and its real TeaCache / @cfg_skip) inside the same loop."""
    import loop_case
    g = loop_case.golden(golden_dir)
    out, decisions = loop_case.run_oracle_loop(g)
    assert decisions == [bool(d) for d in g["decisions"]] and not all(decisions)
    assert _rel(out, torch.from_numpy(g["out"])) < 2e-5


def test_euler_schedule_restatements_agree():
    """Product and oracle restate diffusers' FlowMatchEulerDiscreteScheduler independently (parity unpinned: diffusers
    is absent); they must agree with each other and with the closed form at the ends."""
    from flexam_b200.sampler import flow_match_euler_schedule
    from oracle import sampler_oracle
    for n in (1, 6, 50):
        t1, s1 = flow_match_euler_schedule(n, 5.0)
        t2, s2 = sampler_oracle.euler_schedule(n, 5.0)
        np.testing.assert_allclose(t1, t2, rtol=1e-6)
        np.testing.assert_allclose(s1, s2, rtol=1e-6, atol=1e-9)
        assert s1[0] == 1.0 and s1[-1] == 0.0 and len(s1) == n + 1 and np.all(np.diff(s1) < 0)
    smin = 5.0 * 1e-3 / (1 + 4.0 * 1e-3)                    # the constructor's shifted sigma_min
    np.testing.assert_allclose(s1[-2], 5.0 * smin / (1 + 4.0 * smin), rtol=1e-5)


def test_torch_generator_is_bit_identical_to_numpy():
    """synth.tensor_torch (used to build the 5 B-parameter model on the GPU) == synth.tensor, bit for bit, including
    across the 16 Mi-element chunk boundary and for non-zero means."""
    for name, shape, std, mean in (("w/blocks.3.ffn.0.weight", (1000, 777), 0.018, 0.0), ("in/x", (3, 5, 7), 1.0, 0.0),
                                   ("w/blocks.0.norm3.weight", (3072,), 0.1, 1.0), ("w/big", (4200, 4099), 0.3, 0.0)):
        a = synth.tensor(name, shape, std, mean)
        b = synth.tensor_torch(name, shape, std, mean).numpy()
        assert a.dtype == b.dtype and np.array_equal(a, b), name
    cfg = synth.CONFIGS["tiny"]
    sd_np, sd_t = synth.state_dict(cfg), synth.state_dict_torch(cfg, "cpu")
    assert sd_np.keys() == sd_t.keys() and all(np.array_equal(sd_np[k], sd_t[k].numpy()) for k in sd_np)


def test_rope_tables_with_and_without_riflex_match_the_reference(golden_dir):
    """rope_table() (what the RMSNorm+RoPE kernels read) against the REAL module's `freqs`, default and after
    enable_riflex() (wan_transformer3d_FlexAM.py:56-113, :774-788): fixture rows from oracle/make_golden.py, and the
    mirror class's enable_riflex/disable_riflex switch the engine's table accordingly."""
    from flexam_b200.model import Wan2_2Transformer3DModel_FlexAM, rope_table
    g = np.load(os.path.join(golden_dir, "rope_tables.npz"))
    rows = torch.from_numpy(g["rows"])
    cases = (("plain", None), ("riflex", dict(k=6, L_test=66, L_test_scale=4.886)),
             ("riflex_k4_L49", dict(k=4, L_test=49, L_test_scale=None)))
    for key, rf in cases:
        got = rope_table(128, riflex=rf)[rows].double().numpy()
        np.testing.assert_allclose(got, g[key], rtol=0, atol=6e-8, err_msg=key)      # fp32 rounding of cos / sin
    cfg = synth.CONFIGS["tiny"]
    m = Wan2_2Transformer3DModel_FlexAM(
        model_type="ti2v", patch_size=cfg["patch_size"], text_len=cfg["text_len"], in_dim=cfg["in_dim"], dim=cfg["dim"],
        ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"], out_dim=cfg["out_dim"],
        num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], eps=cfg["eps"], add_ref_conv=True,
        in_dim_ref_conv=cfg["out_dim"], add_cnn_block=True, in_dim_cnn_block=cfg["in_dim_cnn"],
        out_dim_cnn_block=cfg["out_dim_cnn"], device="cpu")
    import flexam_b200.model as fx
    assert m.freqs.dtype == torch.complex128 and tuple(m.freqs.shape) == (1024, 64)      # the reference's attribute
    for call, key in ((m.enable_riflex, "riflex"), (m.disable_riflex, "plain")):
        call()
        np.testing.assert_allclose(torch.view_as_real(m.freqs)[rows].numpy(), g[key], rtol=0, atol=1e-15)
        fx._sync_rope_table(m, m.engine())                    # what every forward does first
        np.testing.assert_allclose(m.engine().freqs[rows].double().numpy(), g[key], rtol=0, atol=6e-8)
    m.freqs = m.freqs.to(device="cpu")                        # callers re-assign it (comfyui nodes.py:324)
    fx._sync_rope_table(m, m.engine())
    assert torch.equal(m.engine().freqs, rope_table(128))


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_installed_module_follows_the_reference_riflex_switch(monkeypatch):
    """install() on a REAL reference module: calling the reference's own enable_riflex()/disable_riflex() afterwards
    (they replace module.freqs) must change the table the native kernels read, and the native result must track the
    reference's torch forward in both states."""
    import cpu_ops_emul
    import flexam_b200.model as fx
    cpu_ops_emul.install(monkeypatch)
    cfg = synth.CONFIGS["tiny"]
    ref = ref_import.build_reference_model(cfg).eval()
    sd = O.to_torch_sd(synth.state_dict(cfg))
    ref.load_state_dict(sd, strict=True)
    inp = synth.inputs(cfg, 3, 4, 8, per_token_t=True, tag="riflex")
    kw = dict(x=torch.from_numpy(inp["x"]), t=torch.from_numpy(inp["t"]),
              context=[torch.from_numpy(c) for c in inp["context"]], seq_len=inp["seq_len"],
              y=torch.from_numpy(inp["y"]), full_ref=torch.from_numpy(inp["full_ref"]),
              additional_control=torch.from_numpy(inp["additional_control"]), density=torch.from_numpy(inp["density"]))
    torch_forward = type(ref).forward
    outs = {}
    with torch.no_grad():
        for state in ("plain", "riflex"):
            if state == "riflex":
                ref.enable_riflex()
            outs[state] = torch_forward(ref, **kw)
        ref.disable_riflex()
    assert _rel(outs["riflex"], outs["plain"]) > 1e-3          # the switch matters for these inputs
    native = fx.install(ref.bfloat16())
    for state in ("plain", "riflex"):
        if state == "riflex":
            native.enable_riflex()                              # the REFERENCE's method: replaces module.freqs
        got = native(**{k: ([c.bfloat16() for c in v] if k == "context" else
                            v.bfloat16() if torch.is_tensor(v) and v.dtype == torch.float32 and k not in ("t", "density")
                            else v) for k, v in kw.items()})
        assert _rel(got, outs[state]) < 1e-2, state


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_installed_reference_module_with_its_own_teacache_and_cfg_skip(monkeypatch, golden_dir):
    """Drop-in check around the REAL module: install() rebinds its forward, the REFERENCE's enable_teacache() /
    enable_cfg_skip() set up the reference's own TeaCache object and counters, and the 6-step sampling loop (kernels
    emulated on the CPU) reproduces the fixture the unmodified module produced: same skip decisions, same latents."""
    import cpu_ops_emul
    import loop_case
    import flexam_b200.model as fx
    from flexam_b200.sampler import DenoiseLoop
    cpu_ops_emul.install(monkeypatch)
    g = loop_case.golden(golden_dir)
    L = loop_case.LOOP
    cfg = synth.CONFIGS[L["config"]]
    ref = ref_import.build_reference_model(cfg).eval()
    ref.load_state_dict(O.to_torch_sd(synth.state_dict(cfg)), strict=True)
    native = fx.install(ref.bfloat16())
    tc = L["teacache"]
    native.enable_teacache(tc["coefficients"], L["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"], offload=False)
    native.enable_cfg_skip(L["cfg_skip_ratio"], L["steps"])
    assert type(native.teacache).__module__.endswith("cache_utils")      # the reference's class, not this package's
    from oracle import make_golden
    lt = make_golden.loop_tensors(cfg, *L["grid"])
    loop = DenoiseLoop(native, lt["latents"], lt["mask"], lt["masked_video_latents"], lt["mask_latents"],
                       lt["control_video_latents"], lt["additional_control"], lt["ref_image_latents"],
                       lt["negative_prompt_embeds"], lt["prompt_embeds"], density=L["density"],
                       guidance_scale=L["guidance"])
    out = loop.run(g["timesteps"], g["sigmas"])
    # the decisions of all steps are computed before the loop from the reference's own TeaCache object (its thresholds,
    # its np.poly1d rescale function, its counters) and must be the ones the unmodified module took
    assert loop.decisions == [bool(d) for d in g["decisions"]]
    assert _rel(out, torch.from_numpy(g["out"])) < 1e-2
    # the per-call form (the reference pipeline's own loop calls forward step by step): same decisions through decide()
    native.enable_teacache(tc["coefficients"], L["steps"], tc["rel_l1_thresh"], tc["num_skip_start_steps"], offload=False)
    native.enable_cfg_skip(L["cfg_skip_ratio"], L["steps"])
    decisions = []
    real_decide = fx._tc.decide
    monkeypatch.setattr(fx._tc, "decide", lambda t, m, c: (lambda r: (decisions.append(bool(r)) if c else None, r)[1])(
        real_decide(t, m, c)))
    monkeypatch.setattr(DenoiseLoop, "teacache_schedule", lambda self, ts: None)
    loop2 = DenoiseLoop(native, lt["latents"], lt["mask"], lt["masked_video_latents"], lt["mask_latents"],
                        lt["control_video_latents"], lt["additional_control"], lt["ref_image_latents"],
                        lt["negative_prompt_embeds"], lt["prompt_embeds"], density=L["density"],
                        guidance_scale=L["guidance"])
    out2 = loop2.run(g["timesteps"], g["sigmas"])
    assert decisions == [bool(d) for d in g["decisions"]] and torch.equal(out2, out)


# ------------------------------------------------------------------------------------------------------
# oracle/library_step.py — the reference's library (cuBLAS / cuDNN / flash-attn) path restated as torch modules; it is
# bench.py's `library_baseline` leg on the B200. Here its module tree is pinned in fp32 on CPU.
# ------------------------------------------------------------------------------------------------------
def _library_model(cfg):
    from oracle.library_step import LibraryStep
    m = LibraryStep(dict(cfg, in_dim_ref_conv=cfg["out_dim"]), backend="sdpa").eval()
    missing, unexpected = m.load_state_dict(O.to_torch_sd(synth.state_dict(cfg)), strict=True)
    assert not missing and not unexpected
    return m


@pytest.mark.parametrize("name", ["tiny_tok", "tiny_sample", "tiny_frac"])
def test_library_step_matches_reference_golden(name, golden_dir):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    F, H, W, per_tok = (int(v) for v in g["meta"])
    cfg = synth.CONFIGS[str(g["config"])]
    inp = synth.inputs(cfg, F, H, W, per_token_t=_tmode(per_tok))
    tt = {k: torch.from_numpy(inp[k]) for k in ("x", "y", "additional_control", "full_ref", "t", "density")}
    out = _library_model(cfg)(tt["x"], tt["t"], [torch.from_numpy(c) for c in inp["context"]], inp["seq_len"], tt["y"],
                              tt["full_ref"], tt["additional_control"], tt["density"])
    assert _rel(out, torch.from_numpy(g["out"])) < 2e-5


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_library_step_has_the_reference_state_dict():
    cfg = synth.CONFIGS["tiny"]
    ref = ref_import.build_reference_model(cfg)
    mine = _library_model(cfg)
    assert {k: tuple(v.shape) for k, v in ref.state_dict().items()} == \
           {k: tuple(v.shape) for k, v in mine.state_dict().items()}
