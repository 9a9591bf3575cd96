"""Diagnostics: per-phase cycle counters of one CTA of the 2-threads-per-row attention pipeline (qimg_set_fmha_trace),
plus a correctness check of the selected mode against torch SDPA.   FT_MODES="4,12,44" python tools/fmha_trace.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vllm_omni_b200 import lib as q  # noqa: E402

dev = torch.device("cuda", 0)
B, H, S_img, T = int(os.environ.get("FT_B", 4)), 24, 4096, 128
S = S_img + T
g = torch.Generator(device=dev).manual_seed(0)
qq, kk, vv = (torch.randn(B, H, S, 128, generator=g, device=dev, dtype=torch.float32).bfloat16() for _ in range(3))
ot = torch.empty(B * T, H * 128, dtype=torch.bfloat16, device=dev)
oi = torch.empty(B * S_img, H * 128, dtype=torch.bfloat16, device=dev)
ref = torch.nn.functional.scaled_dot_product_attention(qq[:1].float(), kk[:1].float(), vv[:1].float())  # [1,H,S,128]
ref = ref.permute(0, 2, 1, 3).reshape(S, H * 128)
trace = torch.zeros(32, dtype=torch.int64, device=dev)
ph = ["loop", "wait_S", "load", "max/xchg", "pingpong", "exp+st", "tail", ""]
names = ["mma_loop", "mma_wait_K", "mma_wait_P1", "mma_wait_V", "mma_wait_P0", "n_kv", "", ""] + \
        [f"sm0_{n}" if n else "" for n in ph] + [f"sm1_{n}" if n else "" for n in ph] + [""] * 8
for mode in [int(m) for m in os.environ.get("FT_MODES", "6,38,14").split(",")]:
    q.set_fmha_mode(mode)
    trace.zero_()
    q.check(q.load().qimg_set_fmha_trace(trace.data_ptr()))
    q.fmha_joint(qq, kk, vv, T, 128 ** -0.5, ot, oi)
    torch.cuda.synchronize()
    q.check(q.load().qimg_set_fmha_trace(None))
    out = torch.cat([ot[:T], oi[:S_img]]).float()
    err = float((out - ref).norm() / ref.norm())
    tr = trace.cpu().tolist()
    n = max(tr[5], 1)
    print(f"mode {mode}: rel_fro vs SDPA fp32 = {err:.3e}")
    print("   per KV tile: " + ", ".join(f"{nm}={v / n:.0f}" for nm, v in zip(names, tr) if nm and nm != "n_kv"))
