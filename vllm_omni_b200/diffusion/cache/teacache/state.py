"""Per-branch TeaCache state (what reference cache/teacache/state.py:9-38 tracks), with device buffers that are
allocated once and overwritten in place instead of re-created every step."""
from __future__ import annotations

import torch


class TeaCacheState:
    __slots__ = ("cnt", "accumulated_rel_l1_distance", "previous_modulated_input", "previous_residual",
                 "previous_residual_encoder", "has_mod", "has_residual")

    def __init__(self) -> None:
        self.previous_modulated_input: torch.Tensor | None = None  # block 0's modulated image stream of the last step
        self.previous_residual: torch.Tensor | None = None          # x_img(after blocks) - x_img(before), last computed step
        self.previous_residual_encoder: torch.Tensor | None = None  # never filled: the text stream does not reach the output
        self.reset()

    def reset(self) -> None:
        """New generation: forget the validity of the buffers (the storage itself is kept for reuse)."""
        self.cnt = 0                              # forwards seen by this branch
        self.accumulated_rel_l1_distance = 0.0
        self.has_mod = False
        self.has_residual = False
