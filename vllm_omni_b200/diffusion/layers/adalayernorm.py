"""AdaLayerNorm CustomOp — drop-in for vllm_omni/diffusion/layers/adalayernorm.py:10-102 whose
`forward_cuda` (there: 4 ATen kernels via forward_native) is ONE fused sm_100a kernel
(qimg_ln_modulate): out = LN(x) * (1 + scale) + shift, returns (out, gate[:, None]).  With `index` (per-token modulation
select of `zero_cond_t` models, :31-54) the same kernel picks each token's modulation row and a gather kernel returns the
per-token gate [B, S, D]."""
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
        B, S, D = x.shape
        mod = mod_params.to(torch.bfloat16).contiguous()  # [B, 3D] = shift | scale | gate   ([2B, 3D] with `index`)
        xb = x.to(torch.bfloat16).contiguous().view(B * S, D)
        if index is None:
            y = qlib.ln_modulate(xb, mod[:, :D], mod[:, D:2 * D], rows_per_batch=S, mod_stride=3 * D, eps=self.eps)
            return y.view(B, S, D), mod[:, 2 * D:].unsqueeze(1)
        # per-token modulation select (reference `preprocess`, :31-54): rows [:B] of mod_params for index == 0, [B:] otherwise
        if mod.shape[0] != 2 * B or tuple(index.shape) != (B, S):
            raise ValueError("with `index` [B, S] the modulation parameters must have 2 * B rows")
        idx = index.to(device=x.device, dtype=torch.int32).contiguous().view(-1)
        y = qlib.ln_modulate_indexed(xb, mod[:, :D], mod[:, D:2 * D], idx, B, rows_per_batch=S, mod_stride=3 * D, eps=self.eps)
        gate = qlib.select_rows(mod[:, 2 * D:], idx, B, rows_per_batch=S)
        return y.view(B, S, D), gate.view(B, S, D)
