"""TEST INFRASTRUCTURE — imports the REAL reference transformer from /root/reference with stub modules.

Only usable in the build container (the reference tree does not travel to the GPU box). It is used by
``oracle/make_golden.py`` to generate the committed fixtures under ``tests/golden/`` and by the CPU tests that
pin ``oracle/flexam_oracle.py`` against the reference when the tree is present. Nothing in the product imports it.

Recipe: ``diffusers`` / ``FlexAM.dist`` are absent, so the five trivial symbols the model file needs
are stubbed, the heavy package ``__init__``s are bypassed with bare package objects, and the real files
``FlexAM/models/{wan_transformer3d_FlexAM,attention_utils,cache_utils}.py`` and ``FlexAM/utils/cfg_optimization.py``
are loaded unmodified.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("FLEXAM_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF_ROOT, "FlexAM", "models", "wan_transformer3d_FlexAM.py"))


def _mod(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _load(name: str, path: str) -> types.ModuleType:
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def import_reference():
    """Returns the reference module object (``FlexAM.models.wan_transformer3d_FlexAM``)."""
    if "FlexAM.models.wan_transformer3d_FlexAM" in sys.modules:
        return sys.modules["FlexAM.models.wan_transformer3d_FlexAM"]
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REF_ROOT}")
    import functools
    import inspect

    import torch

    os.environ.setdefault("VIDEOX_ATTENTION_TYPE", "TORCH_SCALED_DOT")  # flash_attn is importable but CUDA-only

    class _Config(dict):
        __getattr__ = dict.get

    class ConfigMixin:
        config_name = "config.json"

    def register_to_config(init):
        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            self.config = _Config({k: v for k, v in bound.arguments.items() if k != "self"})
            return init(self, *args, **kwargs)
        return wrapper

    class ModelMixin(torch.nn.Module):
        pass

    class FromOriginalModelMixin:
        pass

    class _Logger:
        def __getattr__(self, _):
            return lambda *a, **k: None

    logging = types.SimpleNamespace(get_logger=lambda *_a, **_k: _Logger())

    def is_torch_version(op, ver):
        return True

    _mod("diffusers")
    _mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    _mod("diffusers.loaders")
    _mod("diffusers.loaders.single_file_model", FromOriginalModelMixin=FromOriginalModelMixin)
    _mod("diffusers.models")
    _mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    _mod("diffusers.utils", is_torch_version=is_torch_version, logging=logging)

    pkg = _mod("FlexAM")
    pkg.__path__ = [os.path.join(REF_ROOT, "FlexAM")]
    models = _mod("FlexAM.models")
    models.__path__ = [os.path.join(REF_ROOT, "FlexAM", "models")]

    def _absent(*_a, **_k):
        raise RuntimeError("FlexAM.dist is absent from the reference tree (SURVEY.md F1)")

    _mod("FlexAM.dist", get_sequence_parallel_rank=lambda: 0, get_sequence_parallel_world_size=lambda: 1,
         get_sp_group=_absent, usp_attn_forward=_absent, xFuserLongContextAttention=_absent)
    cfg_opt = _load("FlexAM.utils.cfg_optimization", os.path.join(REF_ROOT, "FlexAM", "utils", "cfg_optimization.py"))
    utils = _mod("FlexAM.utils", cfg_skip=cfg_opt.cfg_skip)
    utils.__path__ = [os.path.join(REF_ROOT, "FlexAM", "utils")]
    return _load("FlexAM.models.wan_transformer3d_FlexAM",
                 os.path.join(REF_ROOT, "FlexAM", "models", "wan_transformer3d_FlexAM.py"))


def build_reference_model(cfg: dict):
    """Instantiate the reference class with the FlexAM yaml's extra kwargs (config/wan2.2/wan_civitai_5b_FlexAM.yaml)."""
    ref = import_reference()
    return ref.Wan2_2Transformer3DModel_FlexAM(
        model_type="ti2v", patch_size=tuple(cfg["patch_size"]), text_len=cfg["text_len"], in_dim=cfg["in_dim"],
        dim=cfg["dim"], ffn_dim=cfg["ffn_dim"], freq_dim=cfg["freq_dim"], text_dim=cfg["text_dim"],
        out_dim=cfg["out_dim"], num_heads=cfg["num_heads"], num_layers=cfg["num_layers"], qk_norm=True,
        cross_attn_norm=True, eps=cfg["eps"], add_ref_conv=True, in_dim_ref_conv=cfg["out_dim"],
        add_cnn_block=True, in_dim_cnn_block=cfg["in_dim_cnn"], out_dim_cnn_block=cfg["out_dim_cnn"])


def build_reference_t5(cfg: dict):
    """The REAL umT5 encoder class (FlexAM/models/wan_text_encoder.py:256-304) built as the FlexAM yaml does
    (shared_pos False, dropout 0)."""
    import_reference()          # installs the diffusers stubs
    name = "FlexAM.models.wan_text_encoder"
    mod = sys.modules.get(name) or _load(name, os.path.join(REF_ROOT, "FlexAM", "models", "wan_text_encoder.py"))
    return mod.WanT5EncoderModel(vocab=cfg["vocab"], dim=cfg["dim"], dim_attn=cfg["dim_attn"], dim_ffn=cfg["dim_ffn"],
                                 num_heads=cfg["num_heads"], num_layers=cfg["num_layers"],
                                 num_buckets=cfg["num_buckets"], shared_pos=False, dropout=0.0)


def build_reference_vae(cfg: dict):
    """The REAL Wan2.2 VAE core (FlexAM/models/wan_vae3_8.py:739-870, AutoencoderKLWan2_2_) with the decoder width of
    ``cfg`` and the encoder width ``cfg["enc_dim"]``."""
    import_reference()
    import torch

    class _Out:
        def __init__(self, sample=None, latent_dist=None):
            self.sample, self.latent_dist = sample, latent_dist
    _mod("diffusers.models.autoencoders")
    _mod("diffusers.models.autoencoders.vae", DecoderOutput=_Out, DiagonalGaussianDistribution=object)
    _mod("diffusers.models.modeling_outputs", AutoencoderKLOutput=_Out)
    _mod("diffusers.utils.accelerate_utils", apply_forward_hook=lambda f: f)
    name = "FlexAM.models.wan_vae3_8"
    mod = sys.modules.get(name) or _load(name, os.path.join(REF_ROOT, "FlexAM", "models", "wan_vae3_8.py"))
    return mod.AutoencoderKLWan2_2_(dim=cfg.get("enc_dim", 16), dec_dim=cfg["dec_dim"], z_dim=cfg["z_dim"],
                                    dim_mult=list(cfg["dim_mult"]),
                                    num_res_blocks=cfg["num_res_blocks"],
                                    temperal_downsample=list(cfg["temperal_downsample"]))
