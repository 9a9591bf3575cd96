"""Measurement tooling (not product code): the reference's op sequence in EAGER PyTorch on the same B200.

SURVEY §8d "GPU reference timing": the reference itself cannot travel to the GPU box, so this times its restatement
(oracle/qwen_image_oracle.py — F.linear / F.layer_norm / F.scaled_dot_product_attention / torch elementwise, the same
ATen kernels the reference's eager path dispatches to: cuBLAS GEMMs, the SDPA flash/cuDNN backend) in bf16 on cuda:0,
on the bench workload (B=4, 1024px, T=128, L=60), and the native engine right after it on identical weights/inputs.
    python tools/eager_baseline.py            # env: EB_LAYERS (60), EB_BATCH (4), EB_RES (1024), EB_ITERS (3)
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qwen_image_oracle as O  # noqa: E402
from vllm_omni_b200 import flops, synthetic  # noqa: E402
from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return out, a.elapsed_time(b) / iters


def main():
    L = int(os.environ.get("EB_LAYERS", "60"))
    B = int(os.environ.get("EB_BATCH", "4"))
    res = int(os.environ.get("EB_RES", "1024"))
    iters = int(os.environ.get("EB_ITERS", "3"))
    T = 128
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    w = dict(synthetic.synthetic_weights(L, seed=0, device=dev, device_generate=True))
    dims = O.DiTDims(num_layers=L)
    lat, txt = (t.to(dev) for t in synthetic.synthetic_inputs(B, res, res, T))
    ts = torch.full((B,), 0.5, dtype=torch.bfloat16, device=dev)
    grid = (1, res // 16, res // 16)
    cpu_rope = O.rope_tables
    O.rope_tables = lambda *a, **k: tuple(t.to(dev) for t in cpu_rope(*a, **k))  # tables are built on the host
    with torch.no_grad():
        out_e, ms_e = timed(lambda: O.model_forward(w, dims, lat, txt, ts, grid), iters)
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            m = QwenImageTransformer2DModel(num_layers=L)
    finally:
        torch.set_default_dtype(torch.float32)
    m.load_weights(w.items())
    del w
    torch.cuda.empty_cache()
    args = (lat, txt, None, ts[:1], [[grid]] * B, [T] * B)
    with torch.no_grad():
        out_n, ms_n = timed(lambda: m(*args, return_dict=False, uniform_timestep=True)[0], iters)
    fl = flops.flops_per_forward(L, res // 16 * (res // 16), T) * B
    print(json.dumps({
        "workload": f"one DiT forward, B={B}, {res}px, T={T}, L={L}, bf16",
        "eager_torch_ms": ms_e, "eager_torch_tflops": fl / ms_e / 1e9,
        "native_ms": ms_n, "native_tflops": fl / ms_n / 1e9, "speedup": ms_e / ms_n,
        "rel_fro_native_vs_eager": O.rel_fro(out_n.cpu(), out_e.cpu()),
        "sdpa_backends": {"flash": torch.backends.cuda.flash_sdp_enabled(), "cudnn": torch.backends.cuda.cudnn_sdp_enabled(),
                          "mem_efficient": torch.backends.cuda.mem_efficient_sdp_enabled()},
        "torch": torch.__version__}))


if __name__ == "__main__":
    main()
