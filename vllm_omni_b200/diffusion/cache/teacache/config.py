"""TeaCache parameters (contract of reference vllm_omni/diffusion/cache/teacache/config.py:9-72): an accumulation
threshold and the five coefficients of the degree-4 polynomial that rescales the step-to-step relative L1 distance of
block 0's modulated input before it is accumulated."""
from __future__ import annotations

import dataclasses

# Highest power first (numpy.poly1d order).  The Qwen-Image values are the ones the reference ships
# (config.py:19-29, taken there from ComfyUI-TeaCache); other transformer classes have no native engine.
QWEN_IMAGE_RESCALE = (-450.0, 280.0, -45.0, 3.2, -0.02)
_MODEL_COEFFICIENTS = {"QwenImageTransformer2DModel": list(QWEN_IMAGE_RESCALE)}


def default_coefficients(transformer_type: str) -> list[float]:
    try:
        return list(_MODEL_COEFFICIENTS[transformer_type])
    except KeyError:
        raise KeyError(f"Cannot find coefficients for {transformer_type}. Supported: {list(_MODEL_COEFFICIENTS)}") from None


@dataclasses.dataclass
class TeaCacheConfig:
    """rel_l1_thresh: reuse the cached residual while the accumulated rescaled distance stays below it (0.2 ~ 1.5x,
    0.4 ~ 1.8x, 0.6 ~ 2x fewer block evaluations per the reference's notes); coefficients: None -> per-model default."""

    rel_l1_thresh: float = 0.2
    coefficients: list[float] | None = None
    transformer_type: str = "QwenImageTransformer2DModel"

    def __post_init__(self) -> None:
        if not self.rel_l1_thresh > 0:
            raise ValueError(f"rel_l1_thresh must be positive, got {self.rel_l1_thresh}")
        if self.coefficients is None:
            self.coefficients = default_coefficients(self.transformer_type)
        self.coefficients = [float(c) for c in self.coefficients]
        if len(self.coefficients) != 5:
            raise ValueError(f"coefficients must contain exactly 5 elements, got {len(self.coefficients)}")
