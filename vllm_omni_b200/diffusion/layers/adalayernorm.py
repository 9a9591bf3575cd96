"""AdaLayerNorm CustomOp — drop-in for vllm_omni/diffusion/layers/adalayernorm.py:10-102 whose
`forward_cuda` (there: 4 ATen kernels via forward_native) is ONE fused sm_100a kernel
(qimg_ln_modulate): out = LN(x) * (1 + scale) + shift, returns (out, gate[:, None])."""
from __future__ import annotations

import torch

from vllm_omni_b200 import lib as qlib
from vllm_omni_b200.diffusion.layers.custom_op import CustomOp


class AdaLayerNorm(CustomOp):
    def __init__(self, hidden_size: int, elementwise_affine: bool = False, eps: float = 1e-6) -> None:
        super().__init__()
        if elementwise_affine:
            raise NotImplementedError("Qwen-Image uses elementwise_affine=False")
        self.eps = eps
        self.elementwise_affine = elementwise_affine
        self.hidden_size = hidden_size

    def forward_cuda(self, x: torch.Tensor, mod_params: torch.Tensor, index: torch.Tensor = None):
        if index is not None:
            raise NotImplementedError("per-token modulation index (zero_cond_t, edit pipelines) is a §8f 'next' item")
        B, S, D = x.shape
        mod = mod_params.to(torch.bfloat16).contiguous()  # [B, 3D] = shift | scale | gate
        y = qlib.ln_modulate(x.to(torch.bfloat16).contiguous().view(B * S, D), mod[:, :D], mod[:, D:2 * D],
                             rows_per_batch=S, mod_stride=3 * D, eps=self.eps)
        return y.view(B, S, D), mod[:, 2 * D:].unsqueeze(1)
