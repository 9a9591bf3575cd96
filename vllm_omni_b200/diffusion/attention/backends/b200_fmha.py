"""B200_FMHA attention backend: `AttentionImpl.forward(q, k, v [B,S,H,hd]) -> [B,S,H,hd]` with the
tcgen05 joint-attention kernel (csrc/qimg_fmha.cuh).  Plugs into the reference selector as a new
`_BACKEND_CONFIG` entry (vllm_omni/diffusion/attention/selector.py:18-32):

    "B200_FMHA": {"module": "vllm_omni_b200.diffusion.attention.backends.b200_fmha", "class": "B200FMHABackend"}

selected with DIFFUSION_ATTENTION_BACKEND=B200_FMHA.  This layer-level entry pays two
layout copies ([B,S,H,hd] <-> head-major); the whole-model engine avoids them by having the QKV GEMM
epilogue write head-major directly."""
from __future__ import annotations

import torch

from vllm_omni_b200 import lib as qlib
from vllm_omni_b200.diffusion.attention.backends.abstract import AttentionBackend, AttentionImpl, AttentionMetadata


class B200FMHAImpl(AttentionImpl):
    def __init__(self, num_heads: int, head_size: int, softmax_scale: float, causal: bool = False,
                 num_kv_heads: int | None = None, prefix: str = "", **extra_impl_args) -> None:
        if head_size != 128:
            raise ValueError("B200_FMHA supports head_size 128 only")
        if causal:
            raise ValueError("B200_FMHA is non-causal (DiT joint attention)")
        if num_kv_heads not in (None, num_heads):
            raise ValueError("B200_FMHA does not implement grouped KV heads")
        self.num_heads, self.softmax_scale = num_heads, softmax_scale

    def forward(self, query, key, value, attn_metadata: AttentionMetadata | None = None):
        if attn_metadata is not None and attn_metadata.attn_mask is not None:
            raise NotImplementedError("attention masks are not used on the Qwen-Image path (sdpa.py:54)")
        if attn_metadata is not None and attn_metadata.joint_query is not None:
            cat = (lambda j, x: torch.cat([j, x], dim=1)) if attn_metadata.joint_strategy == "front" else (
                lambda j, x: torch.cat([x, j], dim=1))
            query, key, value = cat(attn_metadata.joint_query, query), cat(attn_metadata.joint_key, key), cat(
                attn_metadata.joint_value, value)
        B, S, H, hd = query.shape
        q, k, v = (t.to(torch.bfloat16).permute(0, 2, 1, 3).contiguous() for t in (query, key, value))
        # exact pipeline: a per-layer call has no end-of-denoise point at which the fast pipeline's overflow flag could
        # be consulted (the whole-model engine uses the fast one and checks once per denoise)
        out_txt, out_img = qlib.fmha_joint(q, k, v, T=0, softmax_scale=self.softmax_scale, mode=qlib.FMHA_EXACT)
        return out_img.view(B, S, H, hd)


class B200FMHABackend(AttentionBackend):
    accept_output_buffer = False

    @staticmethod
    def get_name() -> str:
        return "B200_FMHA"

    @staticmethod
    def get_impl_cls():
        return B200FMHAImpl

    @staticmethod
    def get_supported_head_sizes() -> list[int]:
        return [128]
