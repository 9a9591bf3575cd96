"""Deterministic synthetic weights / inputs for the Qwen-Image DiT (SURVEY.md §8d
"Synthetic inputs"): every >=1-D matrix N(0, 0.02^2), RMSNorm weights 1.0, biases
N(0, 0.02^2).  There is no network for real checkpoints, so parity and benchmarks run
on random weights at the real shapes.

Names are the reference's parameter names after q/k/v stacking
(qwen_image_transformer.py:804-839; list in SURVEY.md §8c).  Each tensor is seeded from
crc32(name) so the values do not depend on generation order, and a CPU generator is
used unless `device_generate=True` (bench-only: values then differ from the CPU ones).
"""
from __future__ import annotations

import zlib

import torch


def param_shapes(num_layers: int, num_heads: int = 24, head_dim: int = 128, in_channels: int = 64,
                 out_channels: int = 16, patch_size: int = 2, joint_dim: int = 3584) -> dict[str, tuple]:
    D = num_heads * head_dim
    FF = 4 * D
    s: dict[str, tuple] = {
        "time_text_embed.timestep_embedder.linear_1.weight": (D, 256),
        "time_text_embed.timestep_embedder.linear_1.bias": (D,),
        "time_text_embed.timestep_embedder.linear_2.weight": (D, D),
        "time_text_embed.timestep_embedder.linear_2.bias": (D,),
        "txt_norm.weight": (joint_dim,),
        "img_in.weight": (D, in_channels),
        "img_in.bias": (D,),
        "txt_in.weight": (D, joint_dim),
        "txt_in.bias": (D,),
    }
    for i in range(num_layers):
        p = f"transformer_blocks.{i}."
        s.update({
            p + "img_mod.1.weight": (6 * D, D), p + "img_mod.1.bias": (6 * D,),
            p + "txt_mod.1.weight": (6 * D, D), p + "txt_mod.1.bias": (6 * D,),
            p + "attn.to_qkv.weight": (3 * D, D), p + "attn.to_qkv.bias": (3 * D,),
            p + "attn.add_kv_proj.weight": (3 * D, D), p + "attn.add_kv_proj.bias": (3 * D,),
            p + "attn.norm_q.weight": (head_dim,), p + "attn.norm_k.weight": (head_dim,),
            p + "attn.norm_added_q.weight": (head_dim,), p + "attn.norm_added_k.weight": (head_dim,),
            p + "attn.to_out.0.weight": (D, D), p + "attn.to_out.0.bias": (D,),
            p + "attn.to_add_out.weight": (D, D), p + "attn.to_add_out.bias": (D,),
            p + "img_mlp.net.0.proj.weight": (FF, D), p + "img_mlp.net.0.proj.bias": (FF,),
            p + "img_mlp.net.2.weight": (D, FF), p + "img_mlp.net.2.bias": (D,),
            p + "txt_mlp.net.0.proj.weight": (FF, D), p + "txt_mlp.net.0.proj.bias": (FF,),
            p + "txt_mlp.net.2.weight": (D, FF), p + "txt_mlp.net.2.bias": (D,),
        })
    s["norm_out.linear.weight"] = (2 * D, D)
    s["norm_out.linear.bias"] = (2 * D,)
    s["proj_out.weight"] = (patch_size * patch_size * out_channels, D)
    s["proj_out.bias"] = (patch_size * patch_size * out_channels,)
    return s


def _is_norm_weight(name: str) -> bool:
    return name.endswith("norm_q.weight") or name.endswith("norm_k.weight") or name == "txt_norm.weight"


def synthetic_weight(name: str, shape: tuple, seed: int = 0, dtype=torch.bfloat16, device="cpu",
                     device_generate: bool = False, std: float = 0.02, norm_jitter: float = 0.0) -> torch.Tensor:
    s = (zlib.crc32(name.encode()) + 1000003 * seed) & 0x7FFFFFFF
    if device_generate and str(device) != "cpu":
        g = torch.Generator(device=device).manual_seed(s)
        base = torch.randn(shape, generator=g, device=device, dtype=torch.float32)
    else:
        g = torch.Generator().manual_seed(s)
        base = torch.randn(shape, generator=g, dtype=torch.float32)
    if _is_norm_weight(name):
        t = 1.0 + norm_jitter * base
    else:
        t = std * base
    return t.to(dtype).to(device)


def synthetic_weights(num_layers: int, seed: int = 0, dtype=torch.bfloat16, device="cpu", device_generate=False,
                      norm_jitter: float = 0.0, **dims):
    """Iterator of (name, tensor) in checkpoint order — feedable to `load_weights`."""
    for name, shape in param_shapes(num_layers, **dims).items():
        yield name, synthetic_weight(name, shape, seed, dtype, device, device_generate, norm_jitter=norm_jitter)


def split_qkv_checkpoint_names(weights):
    """Turn stacked names back into the q/k/v-separate checkpoint names the reference's
    load_weights consumes (qwen_image_transformer.py:805-815) — used to test the loader."""
    for name, t in weights:
        if ".attn.to_qkv." in name:
            q, k, v = t.chunk(3, dim=0)
            for sub, part in (("to_q", q), ("to_k", k), ("to_v", v)):
                yield name.replace("to_qkv", sub), part
        elif ".attn.add_kv_proj." in name:
            q, k, v = t.chunk(3, dim=0)
            for sub, part in (("add_q_proj", q), ("add_k_proj", k), ("add_v_proj", v)):
                yield name.replace("add_kv_proj", sub), part
        else:
            yield name, t


def synthetic_inputs(batch: int, height_px: int, width_px: int, txt_len: int, joint_dim: int = 3584,
                     dtype=torch.bfloat16, neg: bool = False):
    """Latents (seed 42, randn(B,1,16,H/8,W/8) -> packed [B,S_img,64]) and text embeddings
    (seed 43; negative branch seed 44), as in SURVEY.md §8d."""
    h8, w8 = 2 * (height_px // 16), 2 * (width_px // 16)
    g = torch.Generator().manual_seed(42)
    lat = torch.randn((batch, 1, 16, h8, w8), generator=g, dtype=torch.float32).to(dtype)
    lat = lat.view(batch, 16, h8 // 2, 2, w8 // 2, 2).permute(0, 2, 4, 1, 3, 5).reshape(batch, (h8 // 2) * (w8 // 2), 64)
    g = torch.Generator().manual_seed(43)
    txt = torch.randn((batch, txt_len, joint_dim), generator=g, dtype=torch.float32).to(dtype)
    out = [lat.contiguous(), txt]
    if neg:
        g = torch.Generator().manual_seed(44)
        out.append(torch.randn((batch, txt_len, joint_dim), generator=g, dtype=torch.float32).to(dtype))
    return out


# ----------------------------------------------------------------------------------------------------------------------
# VAE decoder.  Names and shapes are the reference's `AutoencoderKLQwenImage` state dict restricted to the
# decode path (`post_quant_conv.*`, `decoder.*`; autoencoder_kl_qwenimage.py:549-612,707-710).
def vae_decoder_param_shapes(base_dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                             temporal_upsample=(True, True, False), out_channels: int = 3) -> dict[str, tuple]:
    dims = [base_dim * u for u in [dim_mult[-1]] + list(dim_mult[::-1])]  # (:578)
    s: dict[str, tuple] = {
        "post_quant_conv.weight": (z_dim, z_dim, 1, 1, 1), "post_quant_conv.bias": (z_dim,),
        "decoder.conv_in.weight": (dims[0], z_dim, 3, 3, 3), "decoder.conv_in.bias": (dims[0],),
    }

    def resblock(prefix, cin, cout):
        s[f"{prefix}.norm1.gamma"] = (cin, 1, 1, 1)
        s[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3, 3)
        s[f"{prefix}.conv1.bias"] = (cout,)
        s[f"{prefix}.norm2.gamma"] = (cout, 1, 1, 1)
        s[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3, 3)
        s[f"{prefix}.conv2.bias"] = (cout,)
        if cin != cout:
            s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
            s[f"{prefix}.conv_shortcut.bias"] = (cout,)

    d0 = dims[0]
    resblock("decoder.mid_block.resnets.0", d0, d0)
    s["decoder.mid_block.attentions.0.norm.gamma"] = (d0, 1, 1)
    s["decoder.mid_block.attentions.0.to_qkv.weight"] = (3 * d0, d0, 1, 1)
    s["decoder.mid_block.attentions.0.to_qkv.bias"] = (3 * d0,)
    s["decoder.mid_block.attentions.0.proj.weight"] = (d0, d0, 1, 1)
    s["decoder.mid_block.attentions.0.proj.bias"] = (d0,)
    resblock("decoder.mid_block.resnets.1", d0, d0)
    out_dim = d0
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        if i > 0:
            in_dim = in_dim // 2
        cur = in_dim
        for r in range(num_res_blocks + 1):
            resblock(f"decoder.up_blocks.{i}.resnets.{r}", cur, out_dim)
            cur = out_dim
        if i != len(dim_mult) - 1:
            s[f"decoder.up_blocks.{i}.upsamplers.0.resample.1.weight"] = (out_dim // 2, out_dim, 3, 3)
            s[f"decoder.up_blocks.{i}.upsamplers.0.resample.1.bias"] = (out_dim // 2,)
            if temporal_upsample[i]:  # present in the checkpoint, never executed for a single frame (:166-169)
                s[f"decoder.up_blocks.{i}.upsamplers.0.time_conv.weight"] = (out_dim * 2, out_dim, 3, 1, 1)
                s[f"decoder.up_blocks.{i}.upsamplers.0.time_conv.bias"] = (out_dim * 2,)
    s["decoder.norm_out.gamma"] = (out_dim, 1, 1, 1)
    s["decoder.conv_out.weight"] = (out_channels, out_dim, 3, 3, 3)
    s["decoder.conv_out.bias"] = (out_channels,)
    return s


def synthetic_vae_decoder_weights(seed: int = 0, **arch) -> dict[str, torch.Tensor]:
    """fp32 (the dtype the reference loads the VAE in): conv weights N(0, 1/fan_in) so activations stay O(1), biases
    N(0, 0.05^2), RMS gammas 1 + N(0, 0.1^2); each tensor seeded from crc32(name) like the DiT weights."""
    out = {}
    for name, shape in vae_decoder_param_shapes(**arch).items():
        g = torch.Generator(device="cpu")
        g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith(".gamma"):
            t = 1.0 + 0.1 * t
        elif name.endswith(".bias"):
            t = 0.05 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / fan_in) ** 0.5 * (3.0 ** 0.5 if len(shape) == 5 and shape[2] == 3 else 1.0)  # 1 of 3 time taps is live
        out[name] = t
    return out


def vae_encoder_param_shapes(base_dim: int = 96, z_dim: int = 16, dim_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                             temperal_downsample=(False, True, True), in_channels: int = 3) -> dict[str, tuple]:
    """`encoder.*` + `quant_conv.*` of AutoencoderKLQwenImage (autoencoder_kl_qwenimage.py:372-444,706): the encode side the
    edit pipelines run on their condition image (pipeline_qwen_image_edit.py:458-480)."""
    dims = [base_dim * u for u in [1] + list(dim_mult)]  # (:409)
    s: dict[str, tuple] = {"encoder.conv_in.weight": (dims[0], in_channels, 3, 3, 3), "encoder.conv_in.bias": (dims[0],)}

    def resblock(prefix, cin, cout):
        s[f"{prefix}.norm1.gamma"] = (cin, 1, 1, 1)
        s[f"{prefix}.conv1.weight"] = (cout, cin, 3, 3, 3)
        s[f"{prefix}.conv1.bias"] = (cout,)
        s[f"{prefix}.norm2.gamma"] = (cout, 1, 1, 1)
        s[f"{prefix}.conv2.weight"] = (cout, cout, 3, 3, 3)
        s[f"{prefix}.conv2.bias"] = (cout,)
        if cin != cout:
            s[f"{prefix}.conv_shortcut.weight"] = (cout, cin, 1, 1, 1)
            s[f"{prefix}.conv_shortcut.bias"] = (cout,)

    idx = 0  # down_blocks is ONE flat ModuleList of residual blocks and resamplers (:415-427)
    out_dim = dims[0]
    for i, (in_dim, out_dim) in enumerate(zip(dims[:-1], dims[1:])):
        for _ in range(num_res_blocks):
            resblock(f"encoder.down_blocks.{idx}", in_dim, out_dim)
            in_dim = out_dim
            idx += 1
        if i != len(dim_mult) - 1:
            s[f"encoder.down_blocks.{idx}.resample.1.weight"] = (out_dim, out_dim, 3, 3)
            s[f"encoder.down_blocks.{idx}.resample.1.bias"] = (out_dim,)
            if temperal_downsample[i]:  # in the checkpoint, never executed for a single frame (:200-211)
                s[f"encoder.down_blocks.{idx}.time_conv.weight"] = (out_dim, out_dim, 3, 1, 1)
                s[f"encoder.down_blocks.{idx}.time_conv.bias"] = (out_dim,)
            idx += 1
    resblock("encoder.mid_block.resnets.0", out_dim, out_dim)
    s["encoder.mid_block.attentions.0.norm.gamma"] = (out_dim, 1, 1)
    s["encoder.mid_block.attentions.0.to_qkv.weight"] = (3 * out_dim, out_dim, 1, 1)
    s["encoder.mid_block.attentions.0.to_qkv.bias"] = (3 * out_dim,)
    s["encoder.mid_block.attentions.0.proj.weight"] = (out_dim, out_dim, 1, 1)
    s["encoder.mid_block.attentions.0.proj.bias"] = (out_dim,)
    resblock("encoder.mid_block.resnets.1", out_dim, out_dim)
    s["encoder.norm_out.gamma"] = (out_dim, 1, 1, 1)
    s["encoder.conv_out.weight"] = (2 * z_dim, out_dim, 3, 3, 3)
    s["encoder.conv_out.bias"] = (2 * z_dim,)
    s["quant_conv.weight"] = (2 * z_dim, 2 * z_dim, 1, 1, 1)
    s["quant_conv.bias"] = (2 * z_dim,)
    return s


def synthetic_vae_encoder_weights(seed: int = 0, **arch) -> dict[str, torch.Tensor]:
    """Same recipe as `synthetic_vae_decoder_weights` for the encode side."""
    out = {}
    for name, shape in vae_encoder_param_shapes(**arch).items():
        g = torch.Generator(device="cpu")
        g.manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
        t = torch.randn(shape, generator=g, dtype=torch.float32)
        if name.endswith(".gamma"):
            t = 1.0 + 0.1 * t
        elif name.endswith(".bias"):
            t = 0.05 * t
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            t = t * (1.0 / fan_in) ** 0.5 * (3.0 ** 0.5 if len(shape) == 5 and shape[2] == 3 else 1.0)
        out[name] = t
    return out
