"""umT5 text encoder: the oracle restatement against the REAL reference module (committed goldens,
live module when /root/reference is mounted) and the host side of the native encoder with the kernels replaced by their
torch specifications. The kernels themselves are checked on the GPU (tests/test_native_gpu.py)."""
import os

import numpy as np
import pytest
import torch

import cpu_ops_emul
from oracle import ref_import
from oracle import t5_oracle as T


def _rel(a, b):
    return (torch.linalg.vector_norm(a.float() - b.float()) / torch.linalg.vector_norm(b.float())).item()


def _case(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    L, step = int(g["meta"][0]), int(g["meta"][1])
    lens = tuple(int(v) for v in g["meta"][2:])
    cfg = T.T5_CONFIGS[str(g["config"])]
    ids, mask = T.inputs(cfg, L=L, lens=lens)
    return cfg, torch.from_numpy(ids), torch.from_numpy(mask), torch.from_numpy(g["out"]), step


def test_t5_oracle_matches_reference_golden(golden_dir):
    cfg, ids, mask, gold, step = _case(golden_dir, "t5_tiny")
    sd = {k: torch.from_numpy(v) for k, v in T.state_dict(cfg).items()}
    out = T.forward(sd, cfg, ids, mask)
    assert _rel(out[:, ::step], gold) < 2e-5
    emu = T.forward(sd, cfg, ids, mask, policy="bf16")
    assert 1e-4 < _rel(emu, out) < 3e-2          # the all-bf16 module sits ~1e-2 from fp32 on random weights


@pytest.mark.skipif(not ref_import.reference_available(), reason="reference tree not mounted")
def test_t5_oracle_matches_live_reference():
    cfg = T.T5_CONFIGS["tiny"]
    model = ref_import.build_reference_t5(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in T.state_dict(cfg).items()}
    model.load_state_dict(sd, strict=True)
    ids, mask = T.inputs(cfg, L=48, lens=(48, 5), tag="live")
    with torch.no_grad():
        ref = model(torch.from_numpy(ids), torch.from_numpy(mask))[0]
        nomask = model(torch.from_numpy(ids))[0]
    assert _rel(T.forward(sd, cfg, torch.from_numpy(ids), torch.from_numpy(mask)), ref) < 2e-5
    assert _rel(T.forward(sd, cfg, torch.from_numpy(ids), None), nomask) < 2e-5


def test_t5_library_form_equals_the_policy_form_in_fp32(golden_dir):
    """oracle.t5_oracle.forward_library (stock torch ops in the weights' dtype: bench.py's library leg) is the same
    function as the pinned restatement when everything is fp32."""
    cfg, ids, mask, gold, step = _case(golden_dir, "t5_tiny")
    sd = {k: torch.from_numpy(v) for k, v in T.state_dict(cfg).items()}
    assert _rel(T.forward_library(sd, cfg, ids, mask)[:, ::step], gold) < 2e-5


def test_relative_position_buckets_match_between_product_and_oracle():
    from flexam_b200.text_encoder import relative_position_bucket
    for L in (7, 64, 512):
        assert torch.equal(relative_position_bucket(L, L, 32), T.relative_position_bucket(L, L, 32))
    b = T.relative_position_bucket(512, 512, 32)
    assert b.min() == 0 and b.max() == 31 and b[0, 0] == 0 and b[0, 1] == 17 and b[1, 0] == 1   # key after query: +16


def test_t5_param_tree_matches_the_reference_layout():
    from flexam_b200.text_encoder import param_shapes
    cfg = T.T5_CONFIGS["tiny"]
    assert param_shapes(cfg) == {n: tuple(s) for n, s, _, _ in T.param_specs(cfg)}


def test_t5_engine_host_logic_matches_oracle(monkeypatch, golden_dir):
    """flexam_b200.text_encoder.WanT5EncoderModel with emulated kernels vs the bf16-policy oracle and the REAL module's
    fp32 output: launch order, packed q|k|v views, relative-position table indexing, mask handling."""
    from flexam_b200.text_encoder import WanT5EncoderModel
    cpu_ops_emul.install(monkeypatch)
    cfg, ids, mask, gold, step = _case(golden_dir, "t5_tiny")
    m = WanT5EncoderModel(**cfg, device="cpu")
    np_sd = T.state_dict(cfg)
    m.load_state_dict({k: torch.from_numpy(v).bfloat16() for k, v in np_sd.items()}, strict=True)
    out = m(ids, mask)[0]
    assert out.shape == (2, ids.shape[1], cfg["dim"]) and out.dtype == torch.bfloat16
    want = T.forward({k: torch.from_numpy(v) for k, v in np_sd.items()}, cfg, ids, mask, policy="bf16")
    assert _rel(out, want) < 5e-3
    assert _rel(out[:, ::step], gold) < 3e-2
    assert m.engine().launches == 10 * cfg["num_layers"] + 2
    nomask = m(ids)[0]
    assert _rel(nomask, T.forward({k: torch.from_numpy(v) for k, v in np_sd.items()}, cfg, ids, None, policy="bf16")) < 5e-3


def test_t5_engine_edge_shapes_on_one_engine(monkeypatch):
    """One encoder, consecutive calls with different batch sizes and sequence lengths (a single prompt, a one-token
    prompt next to a full-length one, a length that is not a multiple of any tile), each against the oracle."""
    from flexam_b200.text_encoder import WanT5EncoderModel
    cpu_ops_emul.install(monkeypatch)
    cfg = T.T5_CONFIGS["tiny"]
    m = WanT5EncoderModel(**cfg, device="cpu")
    np_sd = T.state_dict(cfg)
    sd = {k: torch.from_numpy(v) for k, v in np_sd.items()}
    m.load_state_dict({k: v.bfloat16() for k, v in sd.items()}, strict=True)
    first = None
    for L, lens in ((24, (7, 24)), (17, (17,)), (40, (1, 40, 13)), (24, (7, 24))):
        ids, mask = T.inputs(cfg, L=L, lens=lens)
        ids, mask = torch.from_numpy(ids), torch.from_numpy(mask)
        out = m(ids, mask)[0]
        want = T.forward(sd, cfg, ids, mask, policy="bf16")
        assert out.shape == want.shape == (len(lens), L, cfg["dim"]) and _rel(out, want) < 5e-3, (L, lens)
        if first is None:
            first = out.clone()
    assert torch.equal(out, first)                       # the first shape again, after the others


def test_t5_from_pretrained_follows_the_reference_contract(tmp_path):
    """.safetensors and torch.save checkpoints, kwargs filtered from the yaml's text_encoder_kwargs (which also carries
    tokenizer entries), non-strict load, dtype cast (reference :306-393)."""
    from safetensors.torch import save_file
    from flexam_b200.text_encoder import WanT5EncoderModel
    cfg = T.T5_CONFIGS["tiny"]
    sd = {k: torch.from_numpy(v) for k, v in T.state_dict(cfg).items()}
    kwargs = dict(cfg, text_encoder_subpath="models_t5.pth", tokenizer_subpath="google/umt5-xxl", text_length=512,
                  shared_pos=False, dropout=0.0)
    save_file({k: v.contiguous() for k, v in sd.items()}, str(tmp_path / "t5.safetensors"))
    extra = dict(sd)
    extra["something.else"] = torch.zeros(3)
    del extra["norm.weight"]
    torch.save(extra, str(tmp_path / "t5.pth"))
    a = WanT5EncoderModel.from_pretrained(str(tmp_path / "t5.safetensors"), additional_kwargs=kwargs)
    assert all(v.dtype == torch.bfloat16 for v in a.state_dict().values())
    assert torch.equal(a.state_dict()["blocks.1.ffn.fc2.weight"], sd["blocks.1.ffn.fc2.weight"].bfloat16())
    b = WanT5EncoderModel.from_pretrained(str(tmp_path / "t5.pth"), additional_kwargs=kwargs, low_cpu_mem_usage=True,
                                          torch_dtype=torch.bfloat16)
    assert torch.equal(b.state_dict()["token_embedding.weight"], sd["token_embedding.weight"].bfloat16())
