"""TEST INFRASTRUCTURE ONLY — runs the UNMODIFIED reference VAE (`AutoencoderKLQwenImage.decode`, vendored at
vllm_omni/diffusion/models/qwen_image/autoencoder_kl_qwenimage.py:667,865; fp32 on the CPU through oracle/ref_shim.py) on
seeded latents and synthetic decoder weights, checks the restatement oracle/vae_oracle.py against it, and stores the
reference outputs in tests/golden/vae_decode_*.pt.  Build container only (the reference does not exist on the GPU box).

Fixtures hold the latents and the reference image; the weights are regenerated from
`vllm_omni_b200.synthetic.synthetic_vae_decoder_weights(seed)` (a checksum is stored to detect RNG drift).
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import ref_shim, vae_oracle  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
# name -> batch, latent grid (h, w), weight seed, latent seed.  `ragged`: no level of the decoder is a multiple of the
# 16 x 8 pixel patch of the native convolution tile, 396 attention positions (K tail of the P*V GEMM); `square`: 256 px.
CASES = {
    "vae_decode_ragged": dict(B=2, grid=(18, 22), wseed=3, zseed=11),
    "vae_decode_256px": dict(B=1, grid=(32, 32), wseed=4, zseed=12),
}


# encode side (edit pipelines' condition image): image size (H, W) with H/8 x W/8 >= 8 x 16
ENC_CASES = {
    "vae_encode_ragged": dict(B=2, size=(144, 176), wseed=13, xseed=21),
}


def weights_checksum(w: dict) -> float:
    return float(sum(v.double().abs().sum() for _, v in sorted(w.items())))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, c in CASES.items():
        W = synthetic.synthetic_vae_decoder_weights(seed=c["wseed"])
        vae = ref_shim.build_reference_vae()
        missing, unexpected = vae.load_state_dict(W, strict=False)
        assert not unexpected and all(k.startswith(("encoder.", "quant_conv.")) for k in missing), (missing, unexpected)
        g = torch.Generator().manual_seed(c["zseed"])
        z = torch.randn(c["B"], 16, 1, *c["grid"], generator=g)
        with torch.no_grad():
            ref = vae.decode(z, return_dict=False)[0]
        ora = vae_oracle.vae_decode(z, W)
        err = (ref - ora).abs().max().item()
        assert err < 1e-4, f"{name}: restatement deviates from the reference by {err}"
        clamped = float((ref.abs() >= 1).float().mean())
        torch.save({"z": z, "image": ref, "wseed": c["wseed"], "weights_checksum": weights_checksum(W), "oracle_max_abs_err": err},
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: image {tuple(ref.shape)} |x| mean {ref.abs().mean():.3f} clamped {clamped:.3f}  oracle-vs-reference max abs {err:.2e}")
    for name, c in ENC_CASES.items():
        W = synthetic.synthetic_vae_encoder_weights(seed=c["wseed"])
        vae = ref_shim.build_reference_vae()
        missing, unexpected = vae.load_state_dict(W, strict=False)
        assert not unexpected and all(k.startswith(("decoder.", "post_quant_conv.")) for k in missing), (missing, unexpected)
        g = torch.Generator().manual_seed(c["xseed"])
        x = torch.rand(c["B"], 3, 1, *c["size"], generator=g) * 2 - 1
        with torch.no_grad():
            dist = vae.encode(x, return_dict=False)[0]  # the reference's own DiagonalGaussianDistribution wrapper is a diffusers class
        ref = dist.parameters
        ora = vae_oracle.vae_encode(x, W)
        err = (ref - ora).abs().max().item()
        assert err < 1e-4, f"{name}: restatement deviates from the reference by {err}"
        torch.save({"x": x, "params": ref, "wseed": c["wseed"], "weights_checksum": weights_checksum(W), "oracle_max_abs_err": err},
                   os.path.join(GOLDEN_DIR, name + ".pt"))
        print(f"{name}: posterior parameters {tuple(ref.shape)} |mean| {ref[:, :16].abs().mean():.3f}  oracle-vs-reference max abs {err:.2e}")


if __name__ == "__main__":
    main()
