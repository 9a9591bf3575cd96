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


# ----------------------------------------------------------------------------------------------------------------------
# VAE decode, TIMING ONLY: the op sequence AutoencoderKLQwenImage._decode dispatches on CUDA for one latent
# frame (autoencoder_kl_qwenimage.py:839-862; fp32 module, cuDNN convolutions with torch's default allow_tf32 = True,
# F.pad + Conv3d per causal convolution, F.scaled_dot_product_attention in the mid block).  Conv3d weights are used as the
# reference holds them ([Co, Ci, 3, 3, 3] on a two-zero-frame padded input), so cuDNN does the work the reference makes it do.
def _vae_causal_conv3d(x, w, b):
    k = w.shape[2]
    pad = (w.shape[4] // 2, w.shape[4] // 2, w.shape[3] // 2, w.shape[3] // 2, 2 * (k // 2), 0)  # (:74-82)
    return F.conv3d(F.pad(x, pad), w, b)


def _vae_rms(x, gamma):
    return F.normalize(x, dim=1) * (x.shape[1] ** 0.5) * gamma


def _vae_resblock(x, W, p):
    h = _vae_causal_conv3d(x, W[p + ".conv_shortcut.weight"], W[p + ".conv_shortcut.bias"]) if (p + ".conv_shortcut.weight") in W else x
    y = F.silu(_vae_rms(x, W[p + ".norm1.gamma"]))
    y = _vae_causal_conv3d(y, W[p + ".conv1.weight"], W[p + ".conv1.bias"])
    y = F.silu(_vae_rms(y, W[p + ".norm2.gamma"]))
    return _vae_causal_conv3d(y, W[p + ".conv2.weight"], W[p + ".conv2.bias"]) + h


def vae_decode_eager(z, W):
    """z [B, 16, 1, h, w] fp32 on the GPU, W the reference-named fp32 state dict on the GPU -> [B, 3, 1, 8h, 8w]."""
    x = _vae_causal_conv3d(z, W["post_quant_conv.weight"], W["post_quant_conv.bias"])
    x = _vae_causal_conv3d(x, W["decoder.conv_in.weight"], W["decoder.conv_in.bias"])
    x = _vae_resblock(x, W, "decoder.mid_block.resnets.0")
    p = "decoder.mid_block.attentions.0"
    B, C, T, h, w = x.shape
    y = x.permute(0, 2, 1, 3, 4).reshape(B * T, C, h, w)
    y = _vae_rms(y, W[p + ".norm.gamma"])
    qkv = F.conv2d(y, W[p + ".to_qkv.weight"], W[p + ".to_qkv.bias"]).reshape(B * T, 1, 3 * C, -1).permute(0, 1, 3, 2).contiguous()
    q, k, v = qkv.chunk(3, dim=-1)
    a = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(B * T, C, h, w)
    a = F.conv2d(a, W[p + ".proj.weight"], W[p + ".proj.bias"])
    x = a.view(B, T, C, h, w).permute(0, 2, 1, 3, 4) + x
    x = _vae_resblock(x, W, "decoder.mid_block.resnets.1")
    i = 0
    while f"decoder.up_blocks.{i}.resnets.0.norm1.gamma" in W:
        r = 0
        while f"decoder.up_blocks.{i}.resnets.{r}.norm1.gamma" in W:
            x = _vae_resblock(x, W, f"decoder.up_blocks.{i}.resnets.{r}")
            r += 1
        up = f"decoder.up_blocks.{i}.upsamplers.0.resample.1"
        if (up + ".weight") in W:
            b_, c_, t_, h_, w_ = x.shape
            y = x.permute(0, 2, 1, 3, 4).reshape(b_ * t_, c_, h_, w_)
            y = F.interpolate(y.float(), scale_factor=(2.0, 2.0), mode="nearest-exact").type_as(y)
            y = F.conv2d(y, W[up + ".weight"], W[up + ".bias"], padding=1)
            x = y.view(b_, t_, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        i += 1
    x = F.silu(_vae_rms(x, W["decoder.norm_out.gamma"]))
    return _vae_causal_conv3d(x, W["decoder.conv_out.weight"], W["decoder.conv_out.bias"]).clamp(-1.0, 1.0)


def vae_encode_eager(x, W):
    """TIMING ONLY: the op sequence `AutoencoderKLQwenImage._encode` dispatches on CUDA for one frame
    (autoencoder_kl_qwenimage.py:793-812): x [B, 3, 1, H, W] fp32 -> posterior parameters [B, 32, 1, H/8, W/8]."""
    h = _vae_causal_conv3d(x, W["encoder.conv_in.weight"], W["encoder.conv_in.bias"])
    i = 0
    while f"encoder.down_blocks.{i}.norm1.gamma" in W or f"encoder.down_blocks.{i}.resample.1.weight" in W:
        p = f"encoder.down_blocks.{i}"
        if (p + ".norm1.gamma") in W:
            h = _vae_resblock(h, W, p)
        else:  # QwenImageResample "downsample2d/3d" on the first frame (:190-199)
            b_, c_, t_, h_, w_ = h.shape
            y = h.permute(0, 2, 1, 3, 4).reshape(b_ * t_, c_, h_, w_)
            y = F.conv2d(F.pad(y, (0, 1, 0, 1)), W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], stride=2)
            h = y.view(b_, t_, y.shape[1], y.shape[2], y.shape[3]).permute(0, 2, 1, 3, 4)
        i += 1
    h = _vae_resblock(h, W, "encoder.mid_block.resnets.0")
    p = "encoder.mid_block.attentions.0"
    B, C, T, hh, ww = h.shape
    y = h.permute(0, 2, 1, 3, 4).reshape(B * T, C, hh, ww)
    y = _vae_rms(y, W[p + ".norm.gamma"])
    qkv = F.conv2d(y, W[p + ".to_qkv.weight"], W[p + ".to_qkv.bias"]).reshape(B * T, 1, 3 * C, -1).permute(0, 1, 3, 2).contiguous()
    q, k, v = qkv.chunk(3, dim=-1)
    a = F.scaled_dot_product_attention(q, k, v).squeeze(1).permute(0, 2, 1).reshape(B * T, C, hh, ww)
    a = F.conv2d(a, W[p + ".proj.weight"], W[p + ".proj.bias"])
    h = a.view(B, T, C, hh, ww).permute(0, 2, 1, 3, 4) + h
    h = _vae_resblock(h, W, "encoder.mid_block.resnets.1")
    h = F.silu(_vae_rms(h, W["encoder.norm_out.gamma"]))
    h = _vae_causal_conv3d(h, W["encoder.conv_out.weight"], W["encoder.conv_out.bias"])
    return _vae_causal_conv3d(h, W["quant_conv.weight"], W["quant_conv.bias"])
