"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED
reference (vllm-omni, /root/reference) on CPU through oracle/ref_shim.py, and checks the
restatement oracle/qwen_image_oracle.py against it while doing so.

Run in the build container only (the reference does not exist on the GPU box):
    python -m oracle.make_golden
Fixtures hold inputs + reference outputs (small); weights are NOT stored — they are
regenerated from `vllm_omni_b200.synthetic` (seeded by crc32(name)); a checksum of the
weights is stored so that RNG drift is detected rather than silently compared.
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import qwen_image_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (num_layers, num_heads, joint_dim, batch, latent grid (h,w), T, weight seed)
CASES = {
    "tiny_L2_H2": dict(L=2, H=2, joint=256, B=2, grid=(8, 6), T=24, seed=1),
    "narrow_L1_H4_ragged": dict(L=1, H=4, joint=192, B=1, grid=(5, 7), T=13, seed=2),
    "fullwidth_L1": dict(L=1, H=24, joint=3584, B=1, grid=(8, 8), T=16, seed=3),
    # depth: the reference's own bf16 path drifts from fp32 with depth (1.16e-2 at L=12), so at depth the
    # criterion is err(native, fp32) <= err(reference-bf16, fp32) + 1e-2 with both numbers recorded here
    "fullwidth_L12": dict(L=12, H=24, joint=3584, B=1, grid=(16, 16), T=32, seed=4),
    # image-edit layout (pipeline_qwen_image_edit.py:602): noisy latents (8x6) followed by one condition image (4x6) on
    # the sequence axis; two RoPE grids, text positions after the larger one
    "tiny_edit_two_grids": dict(L=2, H=2, joint=256, B=2, grid=(8, 6), extra_grids=[(4, 6)], T=24, seed=5),
    # edit-plus layout (pipeline_qwen_image_edit_plus.py:436-464,729-737): TWO condition images of different sizes
    "tiny_edit_three_grids": dict(L=2, H=2, joint=256, B=1, grid=(8, 6), extra_grids=[(4, 6), (3, 5)], T=24, seed=8),
    # the headline shape of BASELINE configs[1] (1024 px: 64x64 latent grid, T=128, D=3072, H=24, S=4224: 33 KV tiles,
    # 17 query-tile pairs, 8-band GEMM raster), depth cut to 2 so the fp32 reference fits the build container
    "fullwidth_1024px_L2": dict(L=2, H=24, joint=3584, B=1, grid=(64, 64), T=128, seed=6),
    # FULL DEPTH (L=60) for criterion (iii) at a reduced width (H=8, D=1024) and grid so that the unmodified reference
    # runs in bf16 AND fp32 on the CPU (60 full-width fp32 blocks = 81 GB do not fit the 62 GB container)
    "narrow_L60_H8": dict(L=60, H=8, joint=1024, B=1, grid=(16, 16), T=32, seed=7),
}


def weights_checksum(w: dict) -> float:
    s = 0.0
    for k in sorted(w):
        s += float(w[k].double().abs().sum())
    return s


def build_case(name: str, c: dict):
    dims = O.DiTDims(num_layers=c["L"], num_heads=c["H"], joint_dim=c["joint"])
    wdict = dict(synthetic.synthetic_weights(c["L"], seed=c["seed"], dtype=torch.bfloat16, norm_jitter=0.1,
                                             num_heads=c["H"], joint_dim=c["joint"]))
    h, w_ = c["grid"]
    grids = [(1, h, w_)] + [(1, a, b) for a, b in c.get("extra_grids", [])]
    S = sum(f * a * b for f, a, b in grids)
    g = torch.Generator().manual_seed(100 + c["seed"])
    hs = torch.randn((c["B"], S, 64), generator=g).bfloat16()
    eh = torch.randn((c["B"], c["T"], c["joint"]), generator=g).bfloat16()
    timestep = (torch.tensor([731.0]).expand(c["B"]).bfloat16() / 1000)
    out = {}
    for dt_name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        model, od = ref_shim.build_reference_model(c["L"], dt, num_attention_heads=c["H"], joint_attention_dim=c["joint"])
        sd = dict(model.named_parameters())
        assert set(sd) == set(wdict), (set(sd) ^ set(wdict))
        with torch.no_grad():
            for k, p in sd.items():
                assert tuple(p.shape) == tuple(wdict[k].shape), k
                p.copy_(wdict[k].to(dt))
        ref = ref_shim.run_reference_model(
            model, od, hidden_states=hs.to(dt), encoder_hidden_states=eh.to(dt),
            encoder_hidden_states_mask=torch.ones(c["B"], c["T"], dtype=torch.long), timestep=timestep.to(dt),
            img_shapes=[list(grids)] * c["B"], txt_seq_lens=[c["T"]] * c["B"])
        mine = O.model_forward(O.cast_weights(wdict, dt), dims, hs.to(dt), eh.to(dt), timestep.to(dt),
                               grids if len(grids) > 1 else grids[0])
        err = O.rel_fro(mine, ref)
        print(f"[{name}] {dt_name}: restatement vs reference rel_fro = {err:.3e}  max|d| = {(mine.float()-ref.float()).abs().max():.3e}")
        assert err < (2e-6 if dt == torch.float32 else 2e-3), "oracle restatement deviates from the reference"
        out[dt_name] = ref.clone()
        del model
    fix = dict(case=c, hidden_states=hs, encoder_hidden_states=eh, timestep=timestep, ref_bf16=out["bf16"],
               ref_fp32=out["fp32"], weights_checksum=weights_checksum(wdict),
               ref_bf16_vs_fp32=O.rel_fro(out["bf16"], out["fp32"]))
    torch.save(fix, os.path.join(GOLDEN_DIR, f"{name}.pt"))
    print(f"[{name}] saved; ref bf16 vs fp32 rel_fro = {fix['ref_bf16_vs_fp32']:.3e}")


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    only = sys.argv[1:]
    for name, c in CASES.items():
        if only and name not in only:
            continue
        build_case(name, c)


if __name__ == "__main__":
    main()
