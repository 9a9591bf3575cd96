"""Step caches on top of the native engine (reference vllm_omni/diffusion/cache/): TeaCache.
cache-dit is a third-party library operating on PyTorch blocks and has no native counterpart here."""
from vllm_omni_b200.diffusion.cache.base import CacheBackend  # noqa: F401
from vllm_omni_b200.diffusion.cache.selector import get_cache_backend  # noqa: F401
