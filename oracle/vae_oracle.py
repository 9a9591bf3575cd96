"""TEST INFRASTRUCTURE ONLY — CPU restatement (fp32 torch ops) of the reference's VAE DECODE of single-frame latents, the
post-step of `QwenImagePipeline.forward` (pipeline_qwen_image.py:736-747 -> AutoencoderKLQwenImage._decode,
vllm_omni/diffusion/models/qwen_image/autoencoder_kl_qwenimage.py:839-862).  Only tests/, smoke() and bench.py's CPU
baseline leg may import this module; the product path is the CUDA decoder (vllm_omni_b200/.../vae_decoder.py).

Parity pinned: `oracle/make_golden_vae.py` runs the UNMODIFIED reference class (through oracle/ref_shim.py) on the same
weights and latents and requires this restatement to agree to fp32 round-off; tests/golden/vae_decode_*.pt hold the
reference outputs.

What the reference does for ONE latent frame (T = 1), restated here without the feature-cache machinery:
 * every QwenImageCausalConv3d(k=3) pads two ZERO frames in front (:78-82, cache empty on the first frame), so only the last
   temporal tap sees data: a 2-D 3x3 convolution with weight[:, :, 2]; k=1 convs (post_quant_conv, conv_shortcut) are 1x1;
 * `upsample3d` skips its time_conv on the first frame (cache slot None -> "Rep", :166-169) and is then identical to
   `upsample2d`: nearest-exact x2 + Conv2d 3x3 (:147-156,190-193);
 * QwenImageRMS_norm = F.normalize(x, dim=channel) * sqrt(C) * gamma (:102-109); SiLU after each norm in the residual
   blocks and before conv_out (:246-279,640-656); one single-head attention over the h*w positions in the mid block
   (:303-331); output clamped to [-1, 1] (:857).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _conv3(x, w, b):
    """x [B, C, h, w]; w the reference's Conv3d weight [Co, Ci, 3, 3, 3] (or Conv2d [Co, Ci, 3, 3])."""
    if w.dim() == 5:
        w = w[:, :, -1]
    return F.conv2d(x, w, b, padding=w.shape[-1] // 2)


def _rms(x, gamma):
    c = x.shape[1]
    return F.normalize(x, dim=1) * (c ** 0.5) * gamma.reshape(1, c, 1, 1)


def _resblock(x, W, p):
    h = _conv3(x, W[p + ".conv_shortcut.weight"], W[p + ".conv_shortcut.bias"]) if (p + ".conv_shortcut.weight") in W else x
    y = F.silu(_rms(x, W[p + ".norm1.gamma"]))
    y = _conv3(y, W[p + ".conv1.weight"], W[p + ".conv1.bias"])
    y = F.silu(_rms(y, W[p + ".norm2.gamma"]))
    y = _conv3(y, W[p + ".conv2.weight"], W[p + ".conv2.bias"])
    return y + h


def _attention(x, W, p):
    B, C, h, w = x.shape
    y = _rms(x, W[p + ".norm.gamma"])
    qkv = F.conv2d(y, W[p + ".to_qkv.weight"], W[p + ".to_qkv.bias"])  # [B, 3C, h, w]
    qkv = qkv.reshape(B, 3 * C, h * w).permute(0, 2, 1)                # [B, hw, 3C]
    q, k, v = qkv.chunk(3, dim=-1)
    a = torch.softmax((q @ k.transpose(1, 2)) * (C ** -0.5), dim=-1) @ v  # SDPA, one head of width C
    a = a.permute(0, 2, 1).reshape(B, C, h, w)
    return F.conv2d(a, W[p + ".proj.weight"], W[p + ".proj.bias"]) + x


def vae_decode(z: torch.Tensor, W: dict, num_up_blocks: int = 4, num_res_blocks: int = 2) -> torch.Tensor:
    """z [B, z_dim, 1, h, w] fp32 (already de-normalised, as `vae.decode` receives it) -> image [B, 3, 1, 8h, 8w] in [-1, 1]."""
    assert z.dim() == 5 and z.shape[2] == 1, "single-frame latents (the image pipelines decode one frame)"
    W = {k: v.float() for k, v in W.items()}
    x = z[:, :, 0].float()
    x = _conv3(x, W["post_quant_conv.weight"], W["post_quant_conv.bias"])
    x = _conv3(x, W["decoder.conv_in.weight"], W["decoder.conv_in.bias"])
    x = _resblock(x, W, "decoder.mid_block.resnets.0")
    x = _attention(x, W, "decoder.mid_block.attentions.0")
    x = _resblock(x, W, "decoder.mid_block.resnets.1")
    for i in range(num_up_blocks):
        for r in range(num_res_blocks + 1):
            x = _resblock(x, W, f"decoder.up_blocks.{i}.resnets.{r}")
        up = f"decoder.up_blocks.{i}.upsamplers.0.resample.1"
        if (up + ".weight") in W:
            x = F.interpolate(x, scale_factor=(2.0, 2.0), mode="nearest-exact")
            x = _conv3(x, W[up + ".weight"], W[up + ".bias"])
    x = F.silu(_rms(x, W["decoder.norm_out.gamma"]))
    x = _conv3(x, W["decoder.conv_out.weight"], W["decoder.conv_out.bias"])
    return x.clamp(-1.0, 1.0).unsqueeze(2)


def vae_encode(x: torch.Tensor, W: dict, num_down_blocks: int = 11) -> torch.Tensor:
    """x [B, 3, 1, H, W] image in [-1, 1] -> posterior parameters [B, 2 z_dim, 1, H/8, W/8] (`AutoencoderKLQwenImage._encode`,
    autoencoder_kl_qwenimage.py:793-812): mean = channels [:z_dim] (what `latent_dist.mode()` / sample_mode="argmax"
    returns, pipeline_qwen_image_edit.py:467), logvar = the rest.

    One frame, restated like the decode: causal 3x3x3 convolutions reduce to 2-D 3x3 with weight[:, :, 2]; both
    `downsample2d` and `downsample3d` are ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2) on the first frame — the time conv
    of `downsample3d` needs a cached previous frame (:200-211) and the first frame only fills the cache."""
    assert x.dim() == 5 and x.shape[2] == 1
    W = {k: v.float() for k, v in W.items()}
    h = _conv3(x[:, :, 0].float(), W["encoder.conv_in.weight"], W["encoder.conv_in.bias"])
    for i in range(num_down_blocks):
        p = f"encoder.down_blocks.{i}"
        if (p + ".norm1.gamma") in W:
            h = _resblock(h, W, p)
        elif (p + ".resample.1.weight") in W:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), W[p + ".resample.1.weight"], W[p + ".resample.1.bias"], stride=2)
    h = _resblock(h, W, "encoder.mid_block.resnets.0")
    h = _attention(h, W, "encoder.mid_block.attentions.0")
    h = _resblock(h, W, "encoder.mid_block.resnets.1")
    h = F.silu(_rms(h, W["encoder.norm_out.gamma"]))
    h = _conv3(h, W["encoder.conv_out.weight"], W["encoder.conv_out.bias"])
    h = _conv3(h, W["quant_conv.weight"], W["quant_conv.bias"])
    return h.unsqueeze(2)
