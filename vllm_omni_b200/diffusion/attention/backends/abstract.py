"""The attention plug-in surface the B200 backend registers through.

It is the contract of the reference's `vllm_omni/diffusion/attention/backends/abstract.py:11-86`, restated: a backend is
a static descriptor (name, implementation class, supported head sizes); an implementation is constructed once per
attention layer and called with `[B, S, H, head_dim]` query / key / value tensors; `AttentionMetadata` carries the
optional text-stream tensors of a joint (text + image) attention and where they go in the joint sequence.
"""
from __future__ import annotations

import abc
import dataclasses

import torch


@dataclasses.dataclass
class AttentionMetadata:
    """Per-call extras.  `joint_*` are the text stream's q/k/v `[B, T, H, head_dim]`; `joint_strategy` says whether
    they are placed in "front" of or at the "rear" of the image tokens (reference ulysses.py:114-121)."""

    attn_mask: torch.Tensor | None = None
    joint_query: torch.Tensor | None = None
    joint_key: torch.Tensor | None = None
    joint_value: torch.Tensor | None = None
    joint_strategy: str = "front"


class AttentionImpl(abc.ABC):
    """One instance per attention layer."""

    @abc.abstractmethod
    def __init__(self, num_heads: int, head_size: int, softmax_scale: float, causal: bool = False,
                 num_kv_heads: int | None = None, prefix: str = "", **extra_impl_args) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                attn_metadata: AttentionMetadata | None = None) -> torch.Tensor:
        """query / key / value `[B, S, H, head_dim]` -> `[B, S, H, head_dim]`."""
        raise NotImplementedError


class AttentionBackend(abc.ABC):
    """Static descriptor looked up by name (`DIFFUSION_ATTENTION_BACKEND`, selector.py)."""

    accept_output_buffer: bool = False

    @staticmethod
    @abc.abstractmethod
    def get_name() -> str:
        raise NotImplementedError

    @staticmethod
    @abc.abstractmethod
    def get_impl_cls() -> type[AttentionImpl]:
        raise NotImplementedError

    @staticmethod
    @abc.abstractmethod
    def get_supported_head_sizes() -> list[int]:
        """An empty list means "any head size"."""
        raise NotImplementedError

    @classmethod
    def supports_head_size(cls, head_size: int) -> bool:
        sizes = cls.get_supported_head_sizes()
        return len(sizes) == 0 or head_size in sizes
