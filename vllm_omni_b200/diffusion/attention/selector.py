"""Backend selection, same mechanism as the reference selector
(vllm_omni/diffusion/attention/selector.py:18-77): DIFFUSION_ATTENTION_BACKEND names a key of
_BACKEND_CONFIG; unknown names raise ValueError.  Here there is exactly one backend and it is
also the default (the reference defaults to SDPA)."""
from __future__ import annotations

import importlib
import os
from functools import cache

_BACKEND_CONFIG = {
    "B200_FMHA": {"module": "vllm_omni_b200.diffusion.attention.backends.b200_fmha", "class": "B200FMHABackend"},
}


def load_backend(name: str):
    cfg = _BACKEND_CONFIG[name]
    return getattr(importlib.import_module(cfg["module"]), cfg["class"])


@cache
def get_attn_backend(head_size: int):
    name = os.environ.get("DIFFUSION_ATTENTION_BACKEND")
    if name is not None:
        if name.upper() not in _BACKEND_CONFIG:
            raise ValueError(f"Invalid attention backend for diffusion: '{name}'. Valid backends are: {list(_BACKEND_CONFIG)}")
        return load_backend(name.upper())
    return load_backend("B200_FMHA")
