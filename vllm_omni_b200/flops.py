"""Algorithmic FLOP / byte counts of the hot path used by bench.py's roofline fields."""
from __future__ import annotations


def flops_per_forward(num_layers: int, s_img: int, t: int, num_heads: int = 24, head_dim: int = 128,
                      in_channels: int = 64, joint_dim: int = 3584) -> float:
    """2*MAC FLOPs of one DiT forward for ONE image: L*(24 S D^2 + 4 S^2 D) + img_in/proj_out + txt_in."""
    D = num_heads * head_dim
    S = s_img + t
    return num_layers * (24.0 * S * D * D + 4.0 * S * S * D) + 2.0 * s_img * in_channels * D * 2 + 2.0 * t * joint_dim * D
