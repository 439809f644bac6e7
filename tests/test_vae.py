"""Wan2.2 VAE decoder (decode half): the oracle restatement against the REAL reference module (committed
golden, live module when /root/reference is mounted) and the host side of the native decoder with the kernels replaced by
their torch specifications. The kernels themselves are checked on the GPU (tests/test_native_gpu.py)."""
import os

import numpy as np
import pytest
import torch

import cpu_ops_emul
from oracle import ref_import
from oracle import vae_oracle as V


def _rel(a, b):
    return (torch.linalg.vector_norm(a.float() - b.float()) / torch.linalg.vector_norm(b.float())).item()


def _case(golden_dir, name="vae_tiny"):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    T, H, W = (int(v) for v in g["meta"])
    cfg = V.VAE_CONFIGS[str(g["config"])]
    return cfg, torch.from_numpy(V.latents(cfg, T, H, W)), V.latent_scale(cfg), torch.from_numpy(g["out"])


def _enc_case(golden_dir, name="vae_tiny"):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    T, H, W = (int(v) for v in g["meta"])
    cfg = V.VAE_CONFIGS[str(g["config"])]
    return cfg, torch.from_numpy(V.video(cfg, 1 + 4 * (T - 1), 16 * H, 16 * W)), V.latent_scale(cfg), torch.from_numpy(g["enc"])


def _full_sd(cfg):
    return {**V.encoder_state_dict(cfg), **V.state_dict(cfg)}


def test_vae_encoder_oracle_matches_reference_golden(golden_dir):
    cfg, x, scale, gold = _enc_case(golden_dir)
    sd = {k: torch.from_numpy(v) for k, v in _full_sd(cfg).items()}
    out = V.encode(sd, cfg, x, scale)
    assert out.shape == gold.shape and _rel(out, gold) < 2e-5


def test_vae_oracle_matches_reference_golden(golden_dir):
    cfg, z, scale, gold = _case(golden_dir)
    sd = {k: torch.from_numpy(v) for k, v in V.state_dict(cfg).items()}
    out = V.decode(sd, cfg, z, scale)
    assert out.shape == gold.shape and _rel(out, gold) < 2e-5
    assert 1e-4 < _rel(V.decode(sd, cfg, z, scale, policy="bf16"), gold) < 3e-2


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_vae_oracle_matches_live_reference():
    """Another latent grid and frame count (5 latent frames -> 17 frames: two cached chunks after the "Rep" one)."""
    cfg = V.VAE_CONFIGS["tiny"]
    model = ref_import.build_reference_vae(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in _full_sd(cfg).items()}
    model.load_state_dict(sd, strict=True)
    z, scale = torch.from_numpy(V.latents(cfg, 5, 2, 4, tag="live")), V.latent_scale(cfg)
    x = torch.from_numpy(V.video(cfg, 13, 32, 48, tag="live"))
    with torch.no_grad():
        ref = model.decode(z, scale).clamp_(-1, 1)
        ref_e = model.encode(x, scale)
    assert ref.shape == (1, 3, 17, 32, 64) and ref_e.shape == (1, 96, 4, 2, 3)
    assert _rel(V.decode(sd, cfg, z, scale), ref) < 2e-5
    assert _rel(V.encode(sd, cfg, x, scale), ref_e) < 2e-5


def test_vae_param_tree_matches_the_reference_layout():
    from flexam_b200.vae import encoder_param_shapes, param_shapes
    for name in ("tiny", "real"):
        cfg = V.VAE_CONFIGS[name]
        assert param_shapes(cfg) == {n: tuple(s) for n, s, _, _ in V.param_specs(cfg)}
        assert encoder_param_shapes(cfg) == {n: tuple(s) for n, s, _, _ in V.encoder_param_specs(cfg)}


def test_vae_decoder_host_logic_matches_oracle(monkeypatch, golden_dir):
    """flexam_b200.vae.AutoencoderKLWan3_8.decode with emulated kernels vs the bf16-policy oracle and the REAL module's
    fp32 output: chunk loop, history grids, tap-major weights, DupUp3D shortcut, frame bookkeeping."""
    from flexam_b200.vae import AutoencoderKLWan3_8
    cpu_ops_emul.install(monkeypatch)
    cfg, z, scale, gold = _case(golden_dir)
    std = 1.0 / scale[1]
    m = AutoencoderKLWan3_8(latent_channels=cfg["z_dim"], c_dim=cfg["enc_dim"], dec_dim=cfg["dec_dim"],
                            dim_mult=cfg["dim_mult"], temperal_downsample=cfg["temperal_downsample"],
                            latents_mean=scale[0], latents_std=std, device="cpu")
    np_sd = _full_sd(cfg)
    m.load_state_dict({"model." + k: torch.from_numpy(v).bfloat16() for k, v in np_sd.items()}, strict=True)
    out = m.decode(z.bfloat16()).sample
    assert out.shape == gold.shape and out.dtype == torch.bfloat16
    want = V.decode({k: torch.from_numpy(v) for k, v in np_sd.items()}, cfg, z, m.scale, policy="bf16")
    r_o, r_g = _rel(out, want), _rel(out, gold)
    print(f"emulated native VAE decode: vs bf16-policy oracle {r_o:.3e}, vs reference golden {r_g:.3e}")
    # two different bf16 emulations of a ~70-op chain (the oracle itself sits 1.2e-2 from fp32 here)
    assert r_o < 2.5e-2 and r_g < 3e-2
    # the encoder half through the same wrapper: vae.encode(x).latent_dist
    _, x, _, gold_e = _enc_case(golden_dir)
    dist = m.encode(x.bfloat16()).latent_dist
    enc = dist.parameters
    assert enc.shape == gold_e.shape and dist.mode().shape[1] == cfg["z_dim"]
    want_e = V.encode({k: torch.from_numpy(v) for k, v in np_sd.items()}, cfg, x, m.scale, policy="bf16")
    r_o, r_g = _rel(enc, want_e), _rel(enc, gold_e)
    print(f"emulated native VAE encode: vs bf16-policy oracle {r_o:.3e}, vs reference golden {r_g:.3e}")
    assert r_o < 2.5e-2 and r_g < 3e-2


def test_vae_single_frame_and_ragged_shapes(monkeypatch):
    """Edge cases of the chunk loops: a single latent frame / a single image (only the first, not time-resampled chunk
    runs), a non-square odd-sized latent grid, and inputs the reference rejects too (frame counts that are not 1 + 4k)."""
    from flexam_b200.lib import FlexamNativeError
    from flexam_b200.vae import AutoencoderKLWan3_8
    cpu_ops_emul.install(monkeypatch)
    cfg = V.VAE_CONFIGS["tiny"]
    scale = V.latent_scale(cfg)
    m = AutoencoderKLWan3_8(latent_channels=cfg["z_dim"], c_dim=cfg["enc_dim"], dec_dim=cfg["dec_dim"],
                            dim_mult=cfg["dim_mult"], temperal_downsample=cfg["temperal_downsample"],
                            latents_mean=scale[0], latents_std=1.0 / scale[1], device="cpu")
    np_sd = _full_sd(cfg)
    sd = {k: torch.from_numpy(v) for k, v in np_sd.items()}
    m.load_state_dict({"model." + k: v.bfloat16() for k, v in sd.items()}, strict=True)
    for T, H, W in ((1, 2, 3), (2, 3, 5)):
        z = torch.from_numpy(V.latents(cfg, T, H, W, tag=f"edge{T}"))
        out = m.decode(z.bfloat16()).sample
        want = V.decode(sd, cfg, z, m.scale, policy="bf16")
        assert out.shape == want.shape == (1, 3, 1 + 4 * (T - 1), 16 * H, 16 * W) and _rel(out, want) < 2.5e-2
    x = torch.from_numpy(V.video(cfg, 1, 32, 48, tag="edge_img"))
    enc = m.encode(x.bfloat16()).latent_dist.parameters
    want = V.encode(sd, cfg, x, m.scale, policy="bf16")
    assert enc.shape == want.shape == (1, 2 * cfg["z_dim"], 1, 2, 3) and _rel(enc, want) < 2.5e-2
    with pytest.raises(FlexamNativeError):
        m.encode(torch.zeros(1, 3, 4, 32, 48, dtype=torch.bfloat16))          # 4 frames: not 1 + 4k
    with pytest.raises(FlexamNativeError):
        m.encode(torch.zeros(1, 3, 5, 40, 48, dtype=torch.bfloat16))          # height not a multiple of 16
    with pytest.raises(FlexamNativeError):
        m.decode(torch.zeros(1, cfg["z_dim"] + 1, 1, 2, 3, dtype=torch.bfloat16))


def test_vae_from_pretrained_and_latent_statistics(tmp_path):
    """The wrapper loader (:1058-1079: inner-model checkpoint, keys prefixed with `model.`) and the Wan2.2 latent
    statistics the wrapper carries by default (:906-1008)."""
    from safetensors.torch import save_file
    from flexam_b200.vae import AutoencoderKLWan3_8, WAN22_LATENTS_MEAN, WAN22_LATENTS_STD
    cfg = V.VAE_CONFIGS["tiny"]
    sd = {k: torch.from_numpy(v) for k, v in _full_sd(cfg).items()}
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "vae.safetensors"))
    kw = dict(latent_channels=cfg["z_dim"], c_dim=cfg["enc_dim"], dec_dim=cfg["dec_dim"], vae_type="AutoencoderKLWan3_8",
              vae_subpath="Wan2.2_VAE.pth")
    m = AutoencoderKLWan3_8.from_pretrained(str(tmp_path / "vae.safetensors"), additional_kwargs=kw)
    got = m.state_dict()
    assert set(got) == {"model." + k for k in sd}
    assert torch.equal(got["model.decoder.head.2.weight"], sd["decoder.head.2.weight"].bfloat16())
    assert len(WAN22_LATENTS_MEAN) == len(WAN22_LATENTS_STD) == 48
    assert torch.allclose(m.scale[0], torch.tensor(WAN22_LATENTS_MEAN)) and \
        torch.allclose(m.scale[1], 1.0 / torch.tensor(WAN22_LATENTS_STD))
