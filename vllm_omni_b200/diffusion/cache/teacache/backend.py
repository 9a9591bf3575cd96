"""reference vllm_omni/diffusion/cache/teacache/backend.py:22-113."""
from __future__ import annotations

import logging
from typing import Any

from vllm_omni_b200.diffusion.cache.base import CacheBackend
from vllm_omni_b200.diffusion.cache.teacache.config import TeaCacheConfig
from vllm_omni_b200.diffusion.cache.teacache.hook import apply_teacache_hook

logger = logging.getLogger(__name__)


class TeaCacheBackend(CacheBackend):
    def enable(self, pipeline: Any) -> None:
        transformer = pipeline.transformer
        transformer_type = transformer.__class__.__name__
        try:
            teacache_config = TeaCacheConfig(transformer_type=transformer_type, rel_l1_thresh=self.config.rel_l1_thresh,
                                             coefficients=self.config.coefficients)
        except Exception as e:
            raise ValueError(f"Invalid TeaCache configuration: {e}. Expected keys: rel_l1_thresh, coefficients (optional). "
                             "transformer_type is automatically extracted from pipeline.transformer.__class__.__name__.")
        apply_teacache_hook(transformer, teacache_config)
        self.enabled = True
        logger.info("TeaCache applied with rel_l1_thresh=%s, transformer_class=%s", teacache_config.rel_l1_thresh, transformer_type)

    def refresh(self, pipeline: Any, num_inference_steps: int, verbose: bool = True) -> None:
        hook = getattr(pipeline.transformer, "_teacache", None)
        if hook is not None:
            hook.reset_state()
