"""AttentionBackend / AttentionImpl / AttentionMetadata — same abstract surface as the reference
(vllm_omni/diffusion/attention/backends/abstract.py:11-86) so the B200 backend registers through
the reference's selector unchanged."""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass

import torch


class AttentionBackend(ABC):
    accept_output_buffer: bool = False

    @staticmethod
    @abstractmethod
    def get_name() -> str: ...

    @staticmethod
    @abstractmethod
    def get_impl_cls() -> type["AttentionImpl"]: ...

    @staticmethod
    @abstractmethod
    def get_supported_head_sizes() -> list[int]: ...

    @classmethod
    def supports_head_size(cls, head_size: int) -> bool:
        s = cls.get_supported_head_sizes()
        return (not s) or head_size in s


@dataclass
class AttentionMetadata:
    attn_mask: torch.Tensor | None = None
    joint_query: torch.Tensor | None = None
    joint_key: torch.Tensor | None = None
    joint_value: torch.Tensor | None = None
    joint_strategy: str = "front"


class AttentionImpl(ABC):
    @abstractmethod
    def __init__(self, num_heads: int, head_size: int, softmax_scale: float, causal: bool = False,
                 num_kv_heads: int | None = None, prefix: str = "", **extra_impl_args) -> None: ...

    @abstractmethod
    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                attn_metadata: AttentionMetadata | None = None) -> torch.Tensor: ...
