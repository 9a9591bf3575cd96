"""Config / output dataclasses of the diffusion runner boundary — field-compatible subsets of
the reference's `vllm_omni/diffusion/data.py` (OmniDiffusionConfig :236-455,
DiffusionParallelConfig :24-91, TransformerConfig :94-117, DiffusionOutput :508-520,
SHUTDOWN_MESSAGE :542).  Only what the DiT denoise path reads is kept; the engine /
entrypoints above the worker stay the reference's own and are out of scope.
"""
from __future__ import annotations

import os

from contextlib import contextmanager
from dataclasses import dataclass, field
from typing import Any

import torch


@dataclass
class DiffusionParallelConfig:
    pipeline_parallel_size: int = 1
    data_parallel_size: int = 1
    tensor_parallel_size: int = 1
    sequence_parallel_size: int | None = None
    ulysses_degree: int = 1
    ring_degree: int = 1
    cfg_parallel_size: int = 1

    def __post_init__(self) -> None:
        if self.sequence_parallel_size is None:
            self.sequence_parallel_size = self.ulysses_degree * self.ring_degree
        for k in ("pipeline_parallel_size", "data_parallel_size", "tensor_parallel_size", "sequence_parallel_size",
                  "ulysses_degree", "ring_degree", "cfg_parallel_size"):
            if getattr(self, k) <= 0:
                raise ValueError(f"{k} must be > 0")
        if self.sequence_parallel_size != self.ulysses_degree * self.ring_degree:
            raise ValueError("sequence_parallel_size must equal ulysses_degree * ring_degree")
        self.world_size = (self.pipeline_parallel_size * self.data_parallel_size * self.tensor_parallel_size *
                           self.ulysses_degree * self.ring_degree * self.cfg_parallel_size)

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "DiffusionParallelConfig":
        if not isinstance(data, dict):
            raise TypeError(f"Expected parallel config dict, got {type(data)!r}")
        return cls(**data)


@dataclass
class TransformerConfig:
    params: dict[str, Any] = field(default_factory=dict)

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "TransformerConfig":
        if not isinstance(data, dict):
            raise TypeError(f"Expected transformer config dict, got {type(data)!r}")
        return cls(params=dict(data))

    def get(self, key: str, default: Any | None = None) -> Any:
        return self.params.get(key, default)

    def __getattr__(self, item: str) -> Any:
        params = object.__getattribute__(self, "params")
        try:
            return params[item]
        except KeyError as exc:
            raise AttributeError(item) from exc


@dataclass
class DiffusionCacheConfig:
    """Cache-adapter parameters (reference data.py:120-200); only the TeaCache ones have a native consumer, unknown keys
    (cache-dit's Fn_compute_blocks, ...) are accepted and kept for interface compatibility."""

    rel_l1_thresh: float = 0.2
    coefficients: list[float] | None = None
    extra: dict[str, Any] = field(default_factory=dict)

    @classmethod
    def from_dict(cls, data: dict[str, Any]) -> "DiffusionCacheConfig":
        if not isinstance(data, dict):
            raise TypeError(f"Expected cache config dict, got {type(data)!r}")
        known = {k: v for k, v in data.items() if k in ("rel_l1_thresh", "coefficients")}
        return cls(**known, extra={k: v for k, v in data.items() if k not in known})


@dataclass
class OmniDiffusionConfig:
    model: str = ""
    model_class_name: str = "QwenImagePipeline"
    dtype: torch.dtype = torch.bfloat16
    tf_model_config: TransformerConfig = field(default_factory=TransformerConfig)
    parallel_config: DiffusionParallelConfig = field(default_factory=DiffusionParallelConfig)
    num_gpus: int | None = None
    master_port: int | None = None
    vae_use_slicing: bool = False
    vae_use_tiling: bool = False
    cache_backend: str | None = "none"  # "tea_cache" (reference data.py:261; env DIFFUSION_CACHE_BACKEND in from_kwargs)
    cache_config: Any = None
    # B200 engine extras (not in the reference): synthetic random weights instead of a checkpoint
    synthetic_weights_seed: int | None = None

    def __post_init__(self):
        if isinstance(self.parallel_config, dict):
            self.parallel_config = DiffusionParallelConfig.from_dict(self.parallel_config)
        if isinstance(self.tf_model_config, dict):
            self.tf_model_config = TransformerConfig.from_dict(self.tf_model_config)
        if self.num_gpus is None:
            self.num_gpus = self.parallel_config.world_size
        if isinstance(self.cache_config, dict):  # reference :438-443
            self.cache_config = DiffusionCacheConfig.from_dict(self.cache_config)
        elif not isinstance(self.cache_config, DiffusionCacheConfig):
            self.cache_config = DiffusionCacheConfig()

    @classmethod
    def from_kwargs(cls, **kwargs) -> "OmniDiffusionConfig":
        if "cache_backend" not in kwargs:  # reference :450-454
            cb = os.environ.get("DIFFUSION_CACHE_BACKEND") or os.environ.get("DIFFUSION_CACHE_ADAPTER")
            kwargs["cache_backend"] = cb.lower() if cb else "none"
        known = {f for f in cls.__dataclass_fields__}
        return cls(**{k: v for k, v in kwargs.items() if k in known})


@dataclass
class DiffusionOutput:
    """Final output after diffusion (reference data.py:508-520)."""

    output: torch.Tensor | None = None
    trajectory_timesteps: list | None = None
    trajectory_latents: torch.Tensor | None = None
    trajectory_decoded: list | None = None
    error: str | None = None


SHUTDOWN_MESSAGE = {"type": "shutdown"}

_current_od_config: OmniDiffusionConfig | None = None


@contextmanager
def set_current_omni_diffusion_config(cfg: OmniDiffusionConfig):
    global _current_od_config
    old = _current_od_config
    _current_od_config = cfg
    try:
        yield
    finally:
        _current_od_config = old


def get_current_omni_diffusion_config() -> OmniDiffusionConfig:
    if _current_od_config is None:
        raise RuntimeError("no current OmniDiffusionConfig")
    return _current_od_config
