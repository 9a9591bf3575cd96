"""Attention pipelines (qimg_set_fmha_mode 4 = exact, 6 = fast, | 8 = 25 % polynomial) at the sweep shapes of
BASELINE configs[4] (512 / 1024 / 2048 px), CUDA events, next to torch SDPA (cuDNN / flash) on the same tensors.
FS_SHAPES="B,S;B,S;..." overrides the shapes (T = 128 text tokens are part of S)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllm_omni_b200 import lib as q  # noqa: E402

dev, bf, H, T = "cuda", torch.bfloat16, 24, 128
shapes = [tuple(int(v) for v in s.split(",")) for s in os.environ.get("FS_SHAPES", "4,4224;1,4224;4,1152;1,16512").split(";")]
modes = [int(m) for m in os.environ.get("FS_MODES", "4,6,12,14").split(",")]
g = torch.Generator(device=dev).manual_seed(0)


def timeit(fn, reps=7):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


print("device", torch.cuda.get_device_name(0))
for B, S in shapes:
    qq, kk, vv = (torch.randn(B, H, S, 128, generator=g, device=dev, dtype=torch.float32).to(bf) for _ in range(3))
    ot = torch.empty(B * T, H * 128, dtype=bf, device=dev)
    oi = torch.empty(B * (S - T), H * 128, dtype=bf, device=dev)
    fl = 4.0 * B * H * S * S * 128
    ref = torch.nn.functional.scaled_dot_product_attention(qq, kk, vv)  # [B,H,S,128]
    ref_img = ref[:, :, T:].permute(0, 2, 1, 3).reshape(B * (S - T), H * 128).float()
    best, med = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv))
    print(f"B={B} S={S}  torch SDPA            best {best:8.3f} ms {fl / best / 1e9:7.1f} TFLOP/s  median {fl / med / 1e9:7.1f}", flush=True)
    for m in modes:
        q.fmha_overflow(reset=True)
        best, med = timeit(lambda: q.fmha_joint(qq, kk, vv, T, 128 ** -0.5, ot, oi, mode=m))
        err = float((oi.float() - ref_img).norm() / ref_img.norm())
        print(f"B={B} S={S}  qimg fmha mode {m:2d}       best {best:8.3f} ms {fl / best / 1e9:7.1f} TFLOP/s  median {fl / med / 1e9:7.1f}"
              f"  rel vs SDPA {err:.2e}  overflow={q.fmha_overflow()}", flush=True)
    del qq, kk, vv, ot, oi, ref, ref_img
