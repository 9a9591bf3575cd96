"""reference vllm_omni/diffusion/cache/selector.py:13-45."""
from __future__ import annotations

from typing import Any

from vllm_omni_b200.diffusion.cache.base import CacheBackend
from vllm_omni_b200.diffusion.data import DiffusionCacheConfig


def get_cache_backend(cache_backend: str | None, cache_config: Any) -> CacheBackend | None:
    if cache_backend is None or cache_backend == "none":
        return None
    if isinstance(cache_config, dict):
        cache_config = DiffusionCacheConfig.from_dict(cache_config)
    elif cache_config is None:
        cache_config = DiffusionCacheConfig()
    if cache_backend == "tea_cache":
        from vllm_omni_b200.diffusion.cache.teacache.backend import TeaCacheBackend
        return TeaCacheBackend(cache_config)
    if cache_backend == "cache_dit":
        raise ValueError("cache_dit drives the third-party cache-dit library over PyTorch blocks; the native engine supports 'tea_cache'")
    raise ValueError(f"Unsupported cache backend: {cache_backend}. Supported: 'tea_cache'")
