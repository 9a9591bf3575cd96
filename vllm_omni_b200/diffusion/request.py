"""OmniDiffusionRequest — the fields of the reference dataclass
(vllm_omni/diffusion/request.py:13-170) that QwenImagePipeline.forward reads, plus
pre-computed embeddings (the reference T2I pipeline ignores req.prompt_embeds,
pipeline_qwen_image.py:602-605 — SURVEY §8f N3; the native runner accepts them so the DiT
stage can run without the in-process Qwen2.5-VL encoder)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

import torch


@dataclass
class OmniDiffusionRequest:
    request_id: str | None = None
    generator: torch.Generator | list[torch.Generator] | None = None
    prompt: str | list[str] | None = None
    negative_prompt: str | list[str] | None = None
    prompt_embeds: torch.Tensor | None = None
    negative_prompt_embeds: torch.Tensor | None = None
    prompt_attention_mask: torch.Tensor | None = None
    negative_attention_mask: torch.Tensor | None = None
    num_outputs_per_prompt: int = 1
    seed: int | None = None
    latents: torch.Tensor | None = None
    height: int | None = None
    width: int | None = None
    num_inference_steps: int = 50
    guidance_scale: float = 1.0
    sigmas: list[float] | None = None
    true_cfg_scale: float | None = None
    output_type: str | None = None
    extra: dict[str, Any] = field(default_factory=dict)
