"""`Attention` module with the reference's constructor/forward (attention/layer.py:17-70), bound to
the B200 backend; no sequence-parallel strategy (Ulysses is the reference's path; the native engine
shards by data / tensor parallelism instead — DESIGN.md)."""
from __future__ import annotations

import torch
import torch.nn as nn

from vllm_omni_b200.diffusion.attention.backends.abstract import AttentionMetadata
from vllm_omni_b200.diffusion.attention.selector import get_attn_backend


class Attention(nn.Module):
    def __init__(self, num_heads: int, head_size: int, causal: bool, softmax_scale: float, num_kv_heads: int | None = None,
                 prefix: str = "", scatter_idx: int = 2, gather_idx: int = 1, use_sync: bool = False):
        super().__init__()
        self.attn_backend = get_attn_backend(-1)
        self.attention = self.attn_backend.get_impl_cls()(num_heads=num_heads, head_size=head_size,
                                                          softmax_scale=softmax_scale, causal=causal,
                                                          num_kv_heads=num_kv_heads)
        self.softmax_scale = softmax_scale

    def forward(self, query: torch.Tensor, key: torch.Tensor, value: torch.Tensor,
                attn_metadata: AttentionMetadata = None) -> torch.Tensor:
        return self.attention.forward(query, key, value, attn_metadata)
