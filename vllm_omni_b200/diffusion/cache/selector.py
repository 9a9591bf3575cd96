"""Cache-backend lookup by name — the role of reference vllm_omni/diffusion/cache/selector.py:13-45 ("tea_cache" /
"cache_dit" / none).  cache-dit is a third-party library that rewrites PyTorch block loops; it has nothing to act on in
the native engine, so only TeaCache resolves."""
from __future__ import annotations

from typing import Any

from vllm_omni_b200.diffusion.cache.base import CacheBackend
from vllm_omni_b200.diffusion.data import DiffusionCacheConfig

_NATIVE = ("tea_cache",)


def get_cache_backend(cache_backend: str | None, cache_config: Any) -> CacheBackend | None:
    name = (cache_backend or "none").lower()
    if name == "none":
        return None
    if name not in _NATIVE:
        hint = " (cache-dit needs PyTorch blocks to patch)" if name == "cache_dit" else ""
        raise ValueError(f"Unsupported cache backend: {cache_backend}{hint}. Supported: {list(_NATIVE)}")
    if cache_config is None:
        cache_config = DiffusionCacheConfig()
    elif isinstance(cache_config, dict):
        cache_config = DiffusionCacheConfig.from_dict(cache_config)
    from vllm_omni_b200.diffusion.cache.teacache.backend import TeaCacheBackend
    return TeaCacheBackend(cache_config)
