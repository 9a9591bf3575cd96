"""TEST INFRASTRUCTURE ONLY — imports the UNMODIFIED reference (vllm-omni) from
/root/reference on CPU so that golden vectors can be generated from the
reference's own code (oracle/make_golden.py) and the restatement in
oracle/qwen_image_oracle.py can be validated against it.

/root/reference only exists in the build container, never on the GPU box:
nothing under tests/ -m gpu, bench.py or smoke() may import this module.

Shims (all verified by running them):
 1. `import vllm_omni` fails against the installed vLLM 0.22 (reference pins 0.12:
    vllm_omni/__init__.py:16 -> patch.py:3).  We pre-register bare namespace
    modules whose __path__ points into /root/reference so the sub-modules on the
    DiT path import cleanly without running the package __init__.
 2. `diffusers` is a third-party, un-vendored dependency (pyproject.toml:35,
    `diffusers>=0.36.0`) that is absent from this image.  A stub provides the five
    symbols the transformer imports (qwen_image_transformer.py:11-16), restating
    the published diffusers semantics:
       FeedForward(dim, dim_out, activation_fn="gelu-approximate")
           = Linear(dim,4dim) -> gelu(tanh) -> Dropout(0) -> Linear(4dim,dim_out)
       Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000)
       TimestepEmbedding(256 -> D): linear_1 -> SiLU -> linear_2
       AdaLayerNormContinuous(D, D, elementwise_affine=False, eps=1e-6):
           emb = linear(silu(c)); scale, shift = chunk(emb, 2); LN(x)*(1+scale)+shift
       Transformer2DModelOutput(sample) supporting out[0]
    These stubbed pieces are "parity unpinned" at the diffusers boundary (the
    reference's tests hold no golden vectors for them) — see DESIGN.md.
 3. A gloo world-size-1 process group + VllmConfig(device=cpu) context so the
    vLLM linear layers construct.
"""
from __future__ import annotations

import math
import os
import sys
import types

REF_ROOT = os.environ.get("QIMG_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vllm_omni"))


def _ns(name: str, path: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


def _install_diffusers_stub():
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    if "diffusers" in sys.modules:
        return

    class GELU(nn.Module):
        def __init__(self, dim_in, dim_out, approximate="none", bias=True):
            super().__init__()
            self.proj = nn.Linear(dim_in, dim_out, bias=bias)
            self.approximate = approximate

        def forward(self, x):
            return F.gelu(self.proj(x), approximate=self.approximate)

    class FeedForward(nn.Module):
        def __init__(self, dim, dim_out=None, mult=4, dropout=0.0, activation_fn="geglu", bias=True):
            super().__init__()
            assert activation_fn == "gelu-approximate"
            inner = int(dim * mult)
            dim_out = dim_out if dim_out is not None else dim
            self.net = nn.ModuleList(
                [GELU(dim, inner, approximate="tanh", bias=bias), nn.Dropout(dropout), nn.Linear(inner, dim_out, bias=bias)]
            )

        def forward(self, x):
            for m in self.net:
                x = m(x)
            return x

    def get_timestep_embedding(timesteps, embedding_dim, flip_sin_to_cos=False, downscale_freq_shift=1.0,
                               scale=1.0, max_period=10000):
        # published diffusers algorithm; the reference restates it in-tree at
        # pipeline_qwen_image.py:135-184
        half_dim = embedding_dim // 2
        exponent = -math.log(max_period) * torch.arange(0, half_dim, dtype=torch.float32, device=timesteps.device)
        exponent = exponent / (half_dim - downscale_freq_shift)
        emb = torch.exp(exponent)
        emb = timesteps[:, None].float() * emb[None, :]
        emb = scale * emb
        emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
        if flip_sin_to_cos:
            emb = torch.cat([emb[:, half_dim:], emb[:, :half_dim]], dim=-1)
        return emb

    class Timesteps(nn.Module):
        def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift, scale=1):
            super().__init__()
            self.num_channels = num_channels
            self.flip_sin_to_cos = flip_sin_to_cos
            self.downscale_freq_shift = downscale_freq_shift
            self.scale = scale

        def forward(self, timesteps):
            return get_timestep_embedding(timesteps, self.num_channels, flip_sin_to_cos=self.flip_sin_to_cos,
                                          downscale_freq_shift=self.downscale_freq_shift, scale=self.scale)

    class TimestepEmbedding(nn.Module):
        def __init__(self, in_channels, time_embed_dim):
            super().__init__()
            self.linear_1 = nn.Linear(in_channels, time_embed_dim, True)
            self.act = nn.SiLU()
            self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim, True)

        def forward(self, sample):
            return self.linear_2(self.act(self.linear_1(sample)))

    class AdaLayerNormContinuous(nn.Module):
        def __init__(self, embedding_dim, conditioning_embedding_dim, elementwise_affine=True, eps=1e-5, bias=True):
            super().__init__()
            self.silu = nn.SiLU()
            self.linear = nn.Linear(conditioning_embedding_dim, embedding_dim * 2, bias=bias)
            self.norm = nn.LayerNorm(embedding_dim, eps, elementwise_affine, bias)

        def forward(self, x, conditioning_embedding):
            emb = self.linear(self.silu(conditioning_embedding).to(x.dtype))
            scale, shift = torch.chunk(emb, 2, dim=1)
            return self.norm(x) * (1 + scale)[:, None, :] + shift[:, None, :]

    class Transformer2DModelOutput:
        def __init__(self, sample):
            self.sample = sample

        def __getitem__(self, i):
            return (self.sample,)[i]

    d = types.ModuleType("diffusers")
    d.__path__ = []
    mods = {
        "diffusers": d,
        "diffusers.models": types.ModuleType("diffusers.models"),
        "diffusers.models.attention": types.ModuleType("diffusers.models.attention"),
        "diffusers.models.embeddings": types.ModuleType("diffusers.models.embeddings"),
        "diffusers.models.modeling_outputs": types.ModuleType("diffusers.models.modeling_outputs"),
        "diffusers.models.normalization": types.ModuleType("diffusers.models.normalization"),
    }
    mods["diffusers.models"].__path__ = []
    mods["diffusers.models.attention"].FeedForward = FeedForward
    mods["diffusers.models.embeddings"].TimestepEmbedding = TimestepEmbedding
    mods["diffusers.models.embeddings"].Timesteps = Timesteps
    mods["diffusers.models.embeddings"].get_timestep_embedding = get_timestep_embedding
    mods["diffusers.models.modeling_outputs"].Transformer2DModelOutput = Transformer2DModelOutput
    mods["diffusers.models.normalization"].AdaLayerNormContinuous = AdaLayerNormContinuous
    sys.modules.update(mods)


def _install_diffusers_vae_stub():
    """diffusers symbols imported by the reference's vendored VAE (autoencoder_kl_qwenimage.py:27-34).  Only what the DECODE
    path touches carries behaviour: `get_activation("silu")` = nn.SiLU() (published diffusers table), `apply_forward_hook`
    = identity (accelerate offload hook), `register_to_config` keeps the constructor untouched; the mixins are empty bases."""
    import torch.nn as nn

    _install_diffusers_stub()
    if "diffusers.models.autoencoders.vae" in sys.modules:
        return

    def mod(name, pkg=False):
        m = sys.modules.get(name) or types.ModuleType(name)
        if pkg:
            m.__path__ = []
        sys.modules[name] = m
        return m

    class ConfigMixin:
        pass

    def register_to_config(init):
        return init

    class FromOriginalModelMixin:
        pass

    class AutoencoderMixin:
        pass

    class ModelMixin(nn.Module):
        pass

    class DecoderOutput:
        def __init__(self, sample):
            self.sample = sample

    class AutoencoderKLOutput:
        def __init__(self, latent_dist):
            self.latent_dist = latent_dist

    class DiagonalGaussianDistribution:  # published diffusers semantics: parameters = (mean, logvar) on the channel axis
        def __init__(self, parameters):
            import torch
            self.parameters = parameters
            self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)

        def mode(self):
            return self.mean

    def get_activation(name):
        if name in ("silu", "swish"):
            return nn.SiLU()
        raise ValueError(name)

    class _Log:
        @staticmethod
        def get_logger(name):
            import logging
            return logging.getLogger(name)

    mod("diffusers.configuration_utils").ConfigMixin = ConfigMixin
    mod("diffusers.configuration_utils").register_to_config = register_to_config
    mod("diffusers.loaders").FromOriginalModelMixin = FromOriginalModelMixin
    mod("diffusers.models.activations").get_activation = get_activation
    mod("diffusers.models.autoencoders", pkg=True)
    v = mod("diffusers.models.autoencoders.vae")
    v.AutoencoderMixin, v.DecoderOutput, v.DiagonalGaussianDistribution = AutoencoderMixin, DecoderOutput, DiagonalGaussianDistribution
    mod("diffusers.models.modeling_outputs").AutoencoderKLOutput = AutoencoderKLOutput
    mod("diffusers.models.modeling_utils").ModelMixin = ModelMixin
    u = mod("diffusers.utils", pkg=True)
    u.logging = _Log
    mod("diffusers.utils.accelerate_utils").apply_forward_hook = lambda f: f


def build_reference_vae(**kwargs):
    """The reference's vendored AutoencoderKLQwenImage (vllm_omni/diffusion/models/qwen_image/autoencoder_kl_qwenimage.py:667),
    unmodified, in fp32 on the CPU (the dtype `from_pretrained` without torch_dtype gives, pipeline_qwen_image.py:267)."""
    import importlib.util

    import torch

    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_diffusers_vae_stub()
    path = os.path.join(REF_ROOT, "vllm_omni", "diffusion", "models", "qwen_image", "autoencoder_kl_qwenimage.py")
    spec = importlib.util.spec_from_file_location("_ref_autoencoder_kl_qwenimage", path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    kwargs.setdefault("dim_mult", [1, 2, 4, 4])  # vae/config.json holds a LIST (the tuple default breaks `[1] + dim_mult`, :409)
    try:
        vae = m.AutoencoderKLQwenImage(**kwargs)
    finally:
        torch.set_default_dtype(prev)
    return vae.eval()


_STATE = {}


def init_reference():
    """Returns (QwenImageTransformer2DModel class, make_config(num_layers, dtype), ctx helpers)."""
    if _STATE:
        return _STATE
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    import torch

    base = os.path.join(REF_ROOT, "vllm_omni")
    _ns("vllm_omni", base)
    _ns("vllm_omni.diffusion.models", os.path.join(base, "diffusion", "models"))
    _ns("vllm_omni.diffusion.models.qwen_image", os.path.join(base, "diffusion", "models", "qwen_image"))
    _install_diffusers_stub()

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    os.environ.setdefault("RANK", "0")
    os.environ.setdefault("LOCAL_RANK", "0")
    os.environ.setdefault("WORLD_SIZE", "1")

    from vllm.config import DeviceConfig, VllmConfig, set_current_vllm_config

    from vllm_omni.diffusion.data import OmniDiffusionConfig, TransformerConfig, set_current_omni_diffusion_config
    from vllm_omni.diffusion.distributed.parallel_state import init_distributed_environment, initialize_model_parallel
    from vllm_omni.diffusion.forward_context import set_forward_context

    vc = VllmConfig(device_config=DeviceConfig(device="cpu"))

    def make_od(num_layers: int, dtype=torch.bfloat16):
        return OmniDiffusionConfig(model="x", dtype=dtype, tf_model_config=TransformerConfig.from_dict({"num_layers": num_layers}))

    od0 = make_od(1)
    with set_current_vllm_config(vc), set_current_omni_diffusion_config(od0):
        if not torch.distributed.is_initialized():
            torch.distributed.init_process_group("gloo", world_size=1, rank=0)
        init_distributed_environment(world_size=1, rank=0, backend="gloo")
        initialize_model_parallel(backend="gloo")
        from vllm_omni.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel

    _STATE.update(
        dict(
            vc=vc,
            make_od=make_od,
            set_current_vllm_config=set_current_vllm_config,
            set_current_omni_diffusion_config=set_current_omni_diffusion_config,
            set_forward_context=set_forward_context,
            Model=QwenImageTransformer2DModel,
        )
    )
    return _STATE


def build_reference_model(num_layers: int, dtype, **dims):
    """Construct the reference QwenImageTransformer2DModel on CPU (`dims` override the
    constructor defaults, e.g. num_attention_heads=2, joint_attention_dim=256)."""
    import torch

    st = init_reference()
    od = st["make_od"](num_layers, dtype)
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        with st["set_current_vllm_config"](st["vc"]), st["set_current_omni_diffusion_config"](od):
            model = st["Model"](od_config=od, **dims)
    finally:
        torch.set_default_dtype(prev)
    model.eval()
    return model, od


def run_reference_model(model, od, **inputs):
    import torch

    st = init_reference()
    with torch.inference_mode(), st["set_forward_context"](vllm_config=st["vc"], omni_diffusion_config=od):
        out = model(return_dict=False, **inputs)
    return out[0]
