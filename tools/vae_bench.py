"""Times the native VAE decode (csrc/qimg_vae.cu) next to the reference's eager CUDA op sequence (baseline/eager_torch.py:
cuDNN TF32 convolutions + SDPA) on the same weights and latents, and reports their difference.  1 GPU, under gpurun.
    VB_SHAPES="B,h,w;..." (latent grid; default 1,128,128;4,128,128 = 1024 px)   VB_ITERS=5"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from baseline import eager_torch  # noqa: E402
from vllm_omni_b200 import lib as q  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402
from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import B200AutoencoderKLQwenImage  # noqa: E402

# conv / GEMM FLOPs of one decode at latent grid (h, w): 2 * pixels * Cin * Cout * taps per layer
def decode_flops(h, w):
    px = h * w
    f = 2 * px * 16 * 384 * 9                                   # conv_in
    f += 4 * 2 * px * 384 * 384 * 9                             # mid res blocks
    f += 2 * px * 384 * 1152 + 2 * px * 384 * 384 + 4 * px * px * 384  # attention: qkv, proj, QK^T + PV
    f += 6 * 2 * px * 384 * 384 * 9                             # up0
    px *= 4
    f += 2 * px * 384 * 192 * 9                                 # up0 resample
    f += 2 * px * 192 * 384 * 9 + 2 * px * 192 * 384 + 5 * 2 * px * 384 * 384 * 9  # up1
    px *= 4
    f += 2 * px * 384 * 192 * 9                                 # up1 resample
    f += 6 * 2 * px * 192 * 192 * 9                             # up2
    px *= 4
    f += 2 * px * 192 * 96 * 9                                  # up2 resample
    f += 6 * 2 * px * 96 * 96 * 9 + 2 * px * 96 * 3 * 9         # up3 + conv_out
    return f


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    dev = "cuda"
    shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("VB_SHAPES", "1,128,128;4,128,128").split(";")]
    iters = int(os.environ.get("VB_ITERS", "5"))
    W = {**synthetic.synthetic_vae_decoder_weights(seed=6), **synthetic.synthetic_vae_encoder_weights(seed=6)}
    Wd = {k: v.to(dev) for k, v in W.items()}
    vae = B200AutoencoderKLQwenImage(W, device=dev)
    for B, h, w in shapes:
        z = torch.randn(B, 16, 1, h, w, generator=torch.Generator().manual_seed(1)).to(dev)
        n0 = q.launch_count()
        out = vae.decode(z, return_dict=False)[0]
        launches = q.launch_count() - n0
        with torch.no_grad():
            ref = eager_torch.vae_decode_eager(z, Wd)
        d = (out - ref).abs()
        t_nat = timed(lambda: vae.decode(z, return_dict=False), iters)
        t_u8 = timed(lambda: vae.decode_to_uint8(z), iters)
        with torch.no_grad():
            t_ref = timed(lambda: eager_torch.vae_decode_eager(z, Wd), iters)
            prev = torch.backends.cudnn.allow_tf32
            torch.backends.cudnn.allow_tf32 = False
            ref32 = eager_torch.vae_decode_eager(z, Wd)
            torch.backends.cudnn.allow_tf32 = prev
        fl = B * decode_flops(h, w)
        print(json.dumps({"B": B, "latent": [h, w], "native_ms": round(t_nat, 3), "native_uint8_ms": round(t_u8, 3),
                          "eager_cudnn_tf32_ms": round(t_ref, 3), "speedup": round(t_ref / t_nat, 2), "launches": launches,
                          "native_tflops": round(fl / t_nat / 1e9, 1), "gflop": round(fl / 1e9, 1),
                          "native_vs_eager_tf32_max_abs": round(d.max().item(), 5),
                          "native_vs_eager_fp32_max_abs": round((out - ref32).abs().max().item(), 5),
                          "eager_tf32_vs_fp32_max_abs": round((ref - ref32).abs().max().item(), 5)}), flush=True)
        # encode side: the edit pipelines' condition image of the same size
        x = (torch.rand(B, 3, 1, 8 * h, 8 * w, generator=torch.Generator().manual_seed(2)) * 2 - 1).to(dev)
        n0 = q.launch_count()
        par = vae.encode(x).latent_dist.parameters
        launches = q.launch_count() - n0
        with torch.no_grad():
            refp = eager_torch.vae_encode_eager(x, Wd)
            t_ref = timed(lambda: eager_torch.vae_encode_eager(x, Wd), iters)
        t_nat = timed(lambda: vae.encode(x), iters)
        print(json.dumps({"encode": True, "B": B, "image": [8 * h, 8 * w], "native_ms": round(t_nat, 3), "eager_cudnn_tf32_ms": round(t_ref, 3),
                          "speedup": round(t_ref / t_nat, 2), "launches": launches,
                          "rel_fro_vs_eager_tf32": round(((par - refp).norm() / refp.norm()).item(), 5)}), flush=True)


if __name__ == "__main__":
    main()
