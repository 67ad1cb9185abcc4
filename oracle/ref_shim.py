"""TEST INFRASTRUCTURE — import shim that lets the UNMODIFIED reference (/root/reference) be imported
in this container, where `diffusers`, `matplotlib` and `seaborn` are not installed.

Used only by oracle/gen_golden.py and the oracle-pinning tests that run where /root/reference exists
(never on the GPU box, never by the product package). It registers stub modules in `sys.modules`
whose *names* satisfy the reference's imports; the only stubs that carry arithmetic are
`Timesteps` / `TimestepEmbedding` (restated from diffusers 0.18.2 `models/embeddings.py`:
sinusoidal embedding with flip_sin_to_cos / freq_shift, then Linear-SiLU-Linear), because the
reference UNet calls them (models/unet_2d_condition.py:284-296, 378-380).
"""
import math
import os
import sys
import types
from collections import OrderedDict

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("RTTI_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def _mod(name):
    m = types.ModuleType(name)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(sys.modules[parent], child, m)
    return m


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels = num_channels
        self.flip_sin_to_cos = flip_sin_to_cos
        self.downscale_freq_shift = downscale_freq_shift

    def forward(self, timesteps):
        half = self.num_channels // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half - self.downscale_freq_shift)
        emb = timesteps[:, None].float() * torch.exp(exponent)[None, :]
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if self.flip_sin_to_cos:
            emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
        return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim, act_fn="silu", out_dim=None, post_act_fn=None, cond_proj_dim=None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

    def forward(self, sample, condition=None):
        return self.linear_2(self.act(self.linear_1(sample)))


def _install_stubs():
    if "diffusers" in sys.modules and getattr(sys.modules["diffusers"], "_rtti_stub", False):
        return
    d = _mod("diffusers")
    d._rtti_stub = True
    for n in ("AutoencoderKL", "PNDMScheduler", "EulerDiscreteScheduler", "DPMSolverMultistepScheduler",
              "StableDiffusionPipeline", "DDIMScheduler"):
        setattr(d, n, type(n, (), {}))

    u = _mod("diffusers.utils")

    class _Logger:
        def __getattr__(self, k):
            return lambda *a, **kw: None

    lg = _mod("diffusers.utils.logging")
    lg.get_logger = lambda *a, **k: _Logger()
    u.logging = lg
    u.deprecate = lambda *a, **k: None
    u.maybe_allow_in_graph = lambda cls: cls
    u.is_torch_version = lambda *a, **k: False
    u.is_accelerate_available = lambda: False
    u.is_accelerate_version = lambda *a, **k: False
    u.is_invisible_watermark_available = lambda: False
    u.randn_tensor = lambda shape, generator=None, device=None, dtype=None: torch.randn(shape, generator=generator, dtype=dtype)
    u.replace_example_docstring = lambda doc: (lambda f: f)

    class BaseOutput(OrderedDict):
        """dict + attribute access; dataclass subclasses populate the dict in __post_init__ (as diffusers does)."""

        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            for k, v in kw.items():
                object.__setattr__(self, k, v)

        def __post_init__(self):
            import dataclasses
            for f in dataclasses.fields(self):
                v = getattr(self, f.name)
                if v is not None:
                    OrderedDict.__setitem__(self, f.name, v)

        def __getitem__(self, k):
            if isinstance(k, int):
                return list(self.values())[k]
            return super().__getitem__(k)

    u.BaseOutput = BaseOutput
    iu = _mod("diffusers.utils.import_utils")
    iu.is_xformers_available = lambda: False

    cu = _mod("diffusers.configuration_utils")

    class _Cfg(dict):
        __getattr__ = dict.get

    class ConfigMixin:
        @property
        def config(self):
            return self._rtti_config

    def register_to_config(init):
        import functools
        import inspect

        @functools.wraps(init)
        def wrapper(self, *args, **kwargs):
            sig = inspect.signature(init)
            bound = sig.bind(self, *args, **kwargs)
            bound.apply_defaults()
            cfg = {k: v for k, v in bound.arguments.items() if k != "self"}
            self._rtti_config = _Cfg(cfg)
            init(self, *args, **kwargs)

        return wrapper

    cu.ConfigMixin = ConfigMixin
    cu.register_to_config = register_to_config
    cu.FrozenDict = dict

    _mod("diffusers.models")
    mu = _mod("diffusers.models.modeling_utils")

    class ModelMixin(nn.Module):
        pass

    mu.ModelMixin = ModelMixin
    ld = _mod("diffusers.loaders")
    for n in ("UNet2DConditionLoadersMixin", "FromSingleFileMixin", "LoraLoaderMixin", "TextualInversionLoaderMixin"):
        setattr(ld, n, type(n, (), {}))
    act = _mod("diffusers.models.activations")

    def get_activation(name):
        return {"silu": nn.SiLU, "swish": nn.SiLU, "mish": nn.Mish, "gelu": nn.GELU, "relu": nn.ReLU}[name]()

    act.get_activation = get_activation
    att = _mod("diffusers.models.attention")
    att.AdaGroupNorm = type("AdaGroupNorm", (nn.Module,), {})
    emb = _mod("diffusers.models.embeddings")
    emb.Timesteps = Timesteps
    emb.TimestepEmbedding = TimestepEmbedding
    for n in ("GaussianFourierProjection", "ImageHintTimeEmbedding", "ImageProjection", "ImageTimeEmbedding",
              "TextImageProjection", "TextImageTimeEmbedding", "TextTimeEmbedding", "PatchEmbed",
              "CombinedTimestepLabelEmbeddings", "ImagePositionalEmbeddings"):
        setattr(emb, n, type(n, (nn.Module,), {}))
    # names the SDXL pipeline file imports at module level
    ip = _mod("diffusers.image_processor")
    ip.VaeImageProcessor = type("VaeImageProcessor", (), {})
    dm = sys.modules["diffusers.models"]
    dm.AutoencoderKL = d.AutoencoderKL
    ap = _mod("diffusers.models.attention_processor")
    for n in ("AttnProcessor2_0", "LoRAAttnProcessor2_0", "LoRAXFormersAttnProcessor", "XFormersAttnProcessor"):
        setattr(ap, n, type(n, (), {}))
    sch = _mod("diffusers.schedulers")
    sch.KarrasDiffusionSchedulers = type("KarrasDiffusionSchedulers", (), {})
    sch.EulerDiscreteScheduler = d.EulerDiscreteScheduler
    _mod("diffusers.pipelines")
    pu = _mod("diffusers.pipelines.pipeline_utils")

    class DiffusionPipeline:
        pass

    pu.DiffusionPipeline = DiffusionPipeline
    _mod("diffusers.pipelines.stable_diffusion_xl")
    sx = sys.modules["diffusers.pipelines.stable_diffusion_xl"]
    sx.StableDiffusionXLPipelineOutput = BaseOutput
    wm = _mod("diffusers.pipelines.stable_diffusion_xl.watermark")
    wm.StableDiffusionXLWatermarker = type("StableDiffusionXLWatermarker", (), {})

    # plotting stack used only for debug JPEGs in utils/attention_utils.py
    class _Anything:
        def __getattr__(self, k):
            return _Anything()

        def __call__(self, *a, **k):
            return _Anything()

        def __iter__(self):
            return iter(())

    for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
        m = _mod(name)
        def _ga(k, _a=_Anything()):  # module-level __getattr__ (PEP 562)
            if k.startswith("__"):
                raise AttributeError(k)
            return _a

        m.__getattr__ = _ga


def import_reference():
    """Returns a namespace with the reference modules (models.*, utils.*) imported unmodified."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self  # `.cuda()` is hard-coded in utils/*.py
    import importlib
    ns = types.SimpleNamespace()
    ns.attention_processor = importlib.import_module("models.attention_processor")
    ns.attention = importlib.import_module("models.attention")
    ns.resnet = importlib.import_module("models.resnet")
    ns.transformer_2d = importlib.import_module("models.transformer_2d")
    ns.unet_2d_blocks = importlib.import_module("models.unet_2d_blocks")
    ns.unet_2d_condition = importlib.import_module("models.unet_2d_condition")
    ns.richtext_utils = importlib.import_module("utils.richtext_utils")
    ns.attention_utils = importlib.import_module("utils.attention_utils")
    ns.attention_utils.plot_attention_maps = lambda *a, **k: None
    ns.region_diffusion = importlib.import_module("models.region_diffusion")
    try:
        ns.region_diffusion_sdxl = importlib.import_module("models.region_diffusion_sdxl")
    except Exception as e:  # the SDXL pipeline file pulls more third-party names; report, do not hide
        ns.region_diffusion_sdxl = None
        ns.region_diffusion_sdxl_error = repr(e)
    return ns
