"""CacheBackend interface — reference vllm_omni/diffusion/cache/base.py:24-90: enable(pipeline), refresh(pipeline,
num_inference_steps, verbose), is_enabled()."""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any

from vllm_omni_b200.diffusion.data import DiffusionCacheConfig


class CacheBackend(ABC):
    def __init__(self, config: DiffusionCacheConfig):
        self.config = config
        self.enabled = False

    @abstractmethod
    def enable(self, pipeline: Any) -> None:
        raise NotImplementedError

    @abstractmethod
    def refresh(self, pipeline: Any, num_inference_steps: int, verbose: bool = True) -> None:
        raise NotImplementedError

    def is_enabled(self) -> bool:
        return self.enabled
