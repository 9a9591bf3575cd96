"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference `AdaLayerNorm.forward_native` WITH a per-token modulation index
(vllm_omni/diffusion/layers/adalayernorm.py:31-54,94-102; the path `zero_cond_t` models take) on CPU and stores inputs and
outputs in tests/golden/adaln_index.pt; checks the restatement `oracle.qwen_image_oracle.ada_layer_norm(..., index)` while
doing so.  Build container only."""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import qwen_image_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "adaln_index.pt")


def main():
    ref_shim.init_reference()
    from vllm_omni.diffusion.layers.adalayernorm import AdaLayerNorm

    g = torch.Generator().manual_seed(77)
    out = {}
    for name, (B, S, D) in {"d256": (2, 37, 256), "d3072": (1, 19, 3072)}.items():
        x = (torch.randn(B, S, D, generator=g) * 2 + 0.3).bfloat16()
        mod = (torch.randn(2 * B, 3 * D, generator=g) * 0.5).bfloat16()
        index = (torch.rand(B, S, generator=g) < 0.4).int()
        index[:, : S // 3] = 0  # the reference's own pattern: first grid 0, condition grids 1 (qwen_image_transformer.py:750-754)
        layer = AdaLayerNorm(D, elementwise_affine=False, eps=1e-6)
        with torch.no_grad():
            y, gate = layer.forward_native(x, mod, index)
            y32, gate32 = layer.forward_native(x.float(), mod.float(), index)
        oy, og = O.ada_layer_norm(x, mod, 1e-6, index)
        assert torch.equal(oy, y) and torch.equal(og.expand_as(gate), gate), name
        out[name] = {"x": x, "mod": mod, "index": index, "y": y, "gate": gate, "y_fp32": y32}
        print(name, tuple(y.shape), tuple(gate.shape), "oracle == reference (bit-exact)")
    torch.save(out, GOLDEN)
    print("saved", GOLDEN)


if __name__ == "__main__":
    main()
