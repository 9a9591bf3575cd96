"""CustomOp — same dispatch surface as the reference (vllm_omni/diffusion/layers/custom_op.py:9-49):
subclasses implement `forward_cuda`; `forward` dispatches to it.  This build targets ONE platform
(CUDA sm_100a), so `dispatch_forward` always returns `forward_cuda` — there is no native/CPU,
HIP or NPU branch to fall back to."""
from __future__ import annotations

from collections.abc import Callable
from typing import Any

import torch.nn as nn


class CustomOp(nn.Module):
    def __init__(self) -> None:
        super().__init__()
        self.is_cuda = True
        self._forward_method = self.dispatch_forward()

    def dispatch_forward(self) -> Callable:
        return self.forward_cuda

    def forward(self, *args, **kwargs) -> Any:
        return self._forward_method(*args, **kwargs)

    def forward_cuda(self, *args, **kwargs):
        raise NotImplementedError

    def forward_native(self, *args, **kwargs):
        raise NotImplementedError("the B200 build has no PyTorch-native path (no CPU fallback by design)")
