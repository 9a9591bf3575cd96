from vllm_omni_b200.diffusion.cache.teacache.backend import TeaCacheBackend  # noqa: F401
from vllm_omni_b200.diffusion.cache.teacache.config import TeaCacheConfig  # noqa: F401
from vllm_omni_b200.diffusion.cache.teacache.hook import TeaCacheHook, apply_teacache_hook  # noqa: F401
from vllm_omni_b200.diffusion.cache.teacache.state import TeaCacheState  # noqa: F401
