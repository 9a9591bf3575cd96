"""TEST INFRASTRUCTURE ONLY — deterministic stand-ins for the two transformers objects the reference pipeline loads for
prompt encoding (`Qwen2Tokenizer`, `Qwen2_5_VLForConditionalGeneration`, pipeline_qwen_image.py:264-272): no tokenizer
files or checkpoints exist offline, and the glue being pinned (chat template, preamble drop, masked extraction, padding,
per-image repeat: :348-434) does not depend on what the encoder computes."""
from __future__ import annotations

import types

import torch
import torch.nn as nn


class _Batch(dict):
    def __getattr__(self, k):
        return self[k]

    def to(self, device):
        return _Batch({k: v.to(device) for k, v in self.items()})


class StubTokenizer:
    """Character-level ids, right-padded to the longest sample (padding=True), truncated to max_length."""

    def __call__(self, texts, max_length=None, padding=True, truncation=True, return_tensors="pt"):
        ids = [[(ord(c) * 7 + 3) % 997 + 1 for c in t][:max_length] for t in texts]
        n = max(len(i) for i in ids)
        input_ids = torch.tensor([i + [0] * (n - len(i)) for i in ids], dtype=torch.long)
        mask = torch.tensor([[1] * len(i) + [0] * (n - len(i)) for i in ids], dtype=torch.long)
        return _Batch(input_ids=input_ids, attention_mask=mask)


class StubTextEncoder(nn.Module):
    """hidden_states[-1][b, t] = E[id] + running masked mean of E over positions <= t (so padding and order matter)."""

    def __init__(self, dim: int = 48, seed: int = 0):
        super().__init__()
        g = torch.Generator().manual_seed(seed)
        self.emb = nn.Parameter(torch.randn(1000, dim, generator=g), requires_grad=False)

    @property
    def dtype(self):
        return self.emb.dtype

    def forward(self, input_ids, attention_mask, output_hidden_states=True):
        e = self.emb[input_ids] * attention_mask[..., None].to(self.emb.dtype)
        run = e.cumsum(dim=1) / attention_mask.cumsum(dim=1).clamp(min=1)[..., None].to(self.emb.dtype)
        return types.SimpleNamespace(hidden_states=(e, e + run))
