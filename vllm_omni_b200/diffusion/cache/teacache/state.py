"""reference vllm_omni/diffusion/cache/teacache/state.py:9-38.  The tensors are device buffers owned by the state and
reused across steps (no per-step allocation)."""
from __future__ import annotations

import torch


class TeaCacheState:
    def __init__(self):
        self.cnt = 0
        self.accumulated_rel_l1_distance = 0.0
        self.previous_modulated_input: torch.Tensor | None = None
        self.previous_residual: torch.Tensor | None = None
        self.previous_residual_encoder: torch.Tensor | None = None  # not kept: the text stream never reaches the output
        self.has_mod = False
        self.has_residual = False

    def reset(self) -> None:
        self.cnt = 0
        self.accumulated_rel_l1_distance = 0.0
        self.has_mod = False
        self.has_residual = False
