"""GPU reference timing (SURVEY.md §8d "GPU reference timing"): the reference's DiT forward as the op sequence its eager
path dispatches on CUDA — cuBLAS GEMMs (F.linear), ATen elementwise / LayerNorm kernels, and the attention backend the
reference selects with DIFFUSION_ATTENTION_BACKEND: torch SDPA (default, backends/sdpa.py:46-66) or flash-attn
(backends/flash_attn.py:44-74, `flash_attn_func`).  /root/reference does not travel to the GPU box, so this is a
restatement of its forward for TIMING ONLY (the parity oracle lives in oracle/ and is never timed on the GPU):

    QwenImageTransformer2DModel.forward      qwen_image_transformer.py:692-802
    QwenImageTransformerBlock.forward        :541-605
    QwenImageCrossAttention.forward          :370-458
    AdaLayerNorm.forward_native              layers/adalayernorm.py:94-102
    apply_rotary_emb_torch (interleaved)     layers/rope.py:13-36   (the reference uses flash-attn's Triton rotary kernel
                                             on CUDA; torch ops here — a handful of extra elementwise launches)

It runs on the SAME parameter tensors as the native model (`dict(model.named_parameters())`), so the comparison is
kernel sequence against kernel sequence on identical weights and inputs.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _rms(x, w, eps):  # vLLM RMSNorm: fp32 normalise -> cast -> * weight
    v = x.float()
    v = v * torch.rsqrt(v.pow(2).mean(-1, keepdim=True) + eps)
    return v.to(x.dtype) * w


def _ada_ln(x, mod, eps):
    shift, scale, gate = mod.chunk(3, dim=-1)
    return F.layer_norm(x, (x.shape[-1],), None, None, eps) * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1), gate.unsqueeze(1)


def _rope(x, cos, sin):
    cos2 = cos.repeat_interleave(2, dim=-1)[:, None, :]
    sin2 = sin.repeat_interleave(2, dim=-1)[:, None, :]
    x1, x2 = x[..., ::2], x[..., 1::2]
    return x * cos2 + torch.stack((-x2, x1), dim=-1).flatten(-2) * sin2


def _attention(q, k, v, scale, backend):
    if backend == "flash_attn":
        from flash_attn import flash_attn_func
        return flash_attn_func(q, k, v, causal=False, softmax_scale=scale)
    o = F.scaled_dot_product_attention(q.permute(0, 2, 1, 3), k.permute(0, 2, 1, 3), v.permute(0, 2, 1, 3), attn_mask=None,
                                       dropout_p=0.0, is_causal=False, scale=scale)
    return o.permute(0, 2, 1, 3)


@torch.no_grad()
def forward(w: dict, num_layers: int, num_heads: int, hidden, enc, timestep, rope, attn_backend: str = "sdpa", eps: float = 1e-6):
    """hidden [B,S_img,64], enc [B,T,joint], timestep [B] (already / 1000), rope = (img_cos, img_sin, txt_cos, txt_sin) in
    the activation dtype on the device -> [B,S_img,64]."""
    dt, H = hidden.dtype, num_heads
    img = F.linear(hidden, w["img_in.weight"], w["img_in.bias"])
    txt = F.linear(_rms(enc, w["txt_norm.weight"], eps), w["txt_in.weight"], w["txt_in.bias"])
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=hidden.device) / half)
    e = (timestep.float() * 1000.0)[:, None] * freqs[None, :]
    temb = torch.cat([torch.cos(e), torch.sin(e)], dim=-1).to(dt)
    p = "time_text_embed.timestep_embedder."
    temb = F.linear(F.silu(F.linear(temb, w[p + "linear_1.weight"], w[p + "linear_1.bias"])), w[p + "linear_2.weight"], w[p + "linear_2.bias"])
    ic, isn, tc, tsn = rope
    T = txt.shape[1]
    for i in range(num_layers):
        b = f"transformer_blocks.{i}."
        im1, im2 = F.linear(F.silu(temb), w[b + "img_mod.1.weight"], w[b + "img_mod.1.bias"]).chunk(2, dim=-1)
        tm1, tm2 = F.linear(F.silu(temb), w[b + "txt_mod.1.weight"], w[b + "txt_mod.1.bias"]).chunk(2, dim=-1)
        xi, gi = _ada_ln(img, im1, eps)
        xt, gt = _ada_ln(txt, tm1, eps)
        iq, ik, iv = (t.unflatten(-1, (H, -1)) for t in F.linear(xi, w[b + "attn.to_qkv.weight"], w[b + "attn.to_qkv.bias"]).chunk(3, dim=-1))
        tq, tk, tv = (t.unflatten(-1, (H, -1)) for t in F.linear(xt, w[b + "attn.add_kv_proj.weight"], w[b + "attn.add_kv_proj.bias"]).chunk(3, dim=-1))
        iq, ik = _rope(_rms(iq, w[b + "attn.norm_q.weight"], eps), ic, isn), _rope(_rms(ik, w[b + "attn.norm_k.weight"], eps), ic, isn)
        tq, tk = _rope(_rms(tq, w[b + "attn.norm_added_q.weight"], eps), tc, tsn), _rope(_rms(tk, w[b + "attn.norm_added_k.weight"], eps), tc, tsn)
        o = _attention(torch.cat([tq, iq], 1), torch.cat([tk, ik], 1), torch.cat([tv, iv], 1), 128 ** -0.5, attn_backend).flatten(2, 3).to(dt)
        img = img + gi * F.linear(o[:, T:], w[b + "attn.to_out.0.weight"], w[b + "attn.to_out.0.bias"])
        txt = txt + gt * F.linear(o[:, :T], w[b + "attn.to_add_out.weight"], w[b + "attn.to_add_out.bias"])
        xi, gi = _ada_ln(img, im2, eps)
        h = F.gelu(F.linear(xi, w[b + "img_mlp.net.0.proj.weight"], w[b + "img_mlp.net.0.proj.bias"]), approximate="tanh")
        img = img + gi * F.linear(h, w[b + "img_mlp.net.2.weight"], w[b + "img_mlp.net.2.bias"])
        xt, gt = _ada_ln(txt, tm2, eps)
        h = F.gelu(F.linear(xt, w[b + "txt_mlp.net.0.proj.weight"], w[b + "txt_mlp.net.0.proj.bias"]), approximate="tanh")
        txt = txt + gt * F.linear(h, w[b + "txt_mlp.net.2.weight"], w[b + "txt_mlp.net.2.bias"])
    emb = F.linear(F.silu(temb), w["norm_out.linear.weight"], w["norm_out.linear.bias"])
    scale, shift = emb.chunk(2, dim=1)
    img = F.layer_norm(img, (img.shape[-1],), None, None, eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    return F.linear(img, w["proj_out.weight"], w["proj_out.bias"])


def time_forward(model, lat, txt, t, grid, T, attn_backend: str, warmup: int = 3, iters: int = 5):
    """CUDA-event time (ms) of one eager forward on the native model's own parameters; None if the backend is unavailable
    on this box (e.g. a flash-attn wheel without sm_100 code)."""
    w = dict(model.named_parameters())
    (ic, isn, tc, tsn), _ = model._rope([[grid]] * lat.shape[0], T, lat.device)
    rope = (ic, isn, tc, tsn)
    ts = t.expand(lat.shape[0]).contiguous()
    try:
        for _ in range(warmup):
            out = forward(w, model.num_layers, model.num_attention_heads, lat, txt, ts, rope, attn_backend)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            out = forward(w, model.num_layers, model.num_attention_heads, lat, txt, ts, rope, attn_backend)
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters, out
    except Exception as exc:  # backend missing / no kernel image for sm_100
        torch.cuda.synchronize()
        return None, repr(exc)[:200]
