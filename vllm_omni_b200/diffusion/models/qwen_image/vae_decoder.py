"""Native VAE decode for the Qwen-Image pipelines.

Stands where the reference puts `AutoencoderKLQwenImage` on the decode side (`self.vae`, pipeline_qwen_image.py:267;
used at :736-747): the same `.dtype`, `.config.{z_dim, latents_mean, latents_std}` and `.decode(z, return_dict=False)[0]`
surface, the same checkpoint keys (`post_quant_conv.*`, `decoder.*` of the VAE sub-folder), but every layer runs on the
sm_100a kernels of csrc/qimg_vae.cu (tcgen05 TF32 implicit-GEMM convolutions on fp32 NHWC activations) instead of ~70
cuDNN / ATen launches over NCHW.  The layer graph below mirrors QwenImageDecoder3d.forward for ONE latent frame
(autoencoder_kl_qwenimage.py:614-662; what "one frame" removes is spelled out in oracle/vae_oracle.py).

There is no PyTorch fallback: without the CUDA library `decode` raises.  torch is used for allocations and the one-time
weight re-layout only.
"""
from __future__ import annotations

from types import SimpleNamespace

import torch

from vllm_omni_b200 import lib as qlib

# AutoencoderKLQwenImage defaults (autoencoder_kl_qwenimage.py:685-690): vae/config.json of Qwen-Image carries the same values
LATENTS_MEAN = [-0.7571, -0.7089, -0.9113, 0.1075, -0.1745, 0.9653, -0.1517, 1.5508, 0.4134, -0.0715, 0.5517, -0.3632, -0.1922,
                -0.9497, 0.2503, -0.2921]
LATENTS_STD = [2.8184, 1.4541, 2.3275, 2.6558, 1.2196, 1.7708, 2.6052, 2.0743, 3.2687, 2.1526, 2.8652, 1.5579, 1.6382, 1.1253,
               2.8251, 1.9160]
SCORES_BYTES_MAX = 4 << 30  # mid-block attention: query rows are processed in bands whose fp32 score matrix fits this


def _pack3x3(w: torch.Tensor, cin_pad: int | None = None) -> torch.Tensor:
    """Conv3d [Co, Ci, 3, 3, 3] (last temporal tap: the only one that meets data on the first frame) or Conv2d [Co, Ci, 3, 3]
    -> [Co, 9 * Ci_pad] with column (3 * ky + kx) * Ci_pad + ci."""
    if w.dim() == 5:
        w = w[:, :, -1]
    co, ci = w.shape[0], w.shape[1]
    cp = ci if cin_pad is None else cin_pad
    out = torch.zeros((co, 3, 3, cp), dtype=torch.float32, device=w.device)
    out[..., :ci] = w.permute(0, 2, 3, 1)
    return out.reshape(co, 9 * cp).contiguous()


def _pack1x1(w: torch.Tensor) -> torch.Tensor:
    return w.reshape(w.shape[0], w.shape[1]).contiguous()


class B200VaeDecoder:
    PREFIXES = ("decoder.", "post_quant_conv.")

    def __init__(self, state_dict: dict, device="cuda", latents_mean=None, latents_std=None, z_dim: int = 16):
        self.device = torch.device(device)
        self.dtype = torch.float32  # what the pipeline casts the latents to (:737); the reference loads the VAE in fp32
        self.config = SimpleNamespace(z_dim=z_dim, latents_mean=list(latents_mean or LATENTS_MEAN),
                                      latents_std=list(latents_std or LATENTS_STD))
        self.temperal_downsample = [False, True, True]  # read by the edit pipelines for vae_scale_factor (:235)
        sd = {k: v.detach().to(self.device, torch.float32) for k, v in state_dict.items() if k.startswith(self.PREFIXES)}
        if "decoder.conv_in.weight" not in sd:
            raise ValueError("state dict holds no `decoder.*` / `post_quant_conv.*` tensors (AutoencoderKLQwenImage checkpoint keys)")
        self.w: dict[str, torch.Tensor] = {}
        for k, v in sd.items():
            if ".time_conv." in k:
                continue  # never executed for a single frame (autoencoder_kl_qwenimage.py:166-169,200-211)
            if k.endswith(".gamma"):
                self.w[k] = v.reshape(-1).contiguous()
            elif k.endswith(".bias"):
                self.w[k] = v.contiguous()
            elif k in ("decoder.conv_in.weight", "encoder.conv_in.weight"):
                self.w[k] = _pack3x3(v, cin_pad=32)
            elif k == "decoder.conv_out.weight":
                self.w[k] = v[:, :, -1].permute(0, 2, 3, 1).contiguous()  # [3, ky, kx, C]
            elif v.dim() >= 4 and v.shape[-1] == 3:
                self.w[k] = _pack3x3(v)
            else:
                self.w[k] = _pack1x1(v)
        self.num_up_blocks = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("decoder.up_blocks."))
        self.num_res = 1 + max(int(k.split(".")[4]) for k in sd if k.startswith("decoder.up_blocks.0.resnets."))

    # -- layers ---------------------------------------------------------------------------------------------------------
    def _conv3(self, x, name, res=None):
        w = self.w[name + ".weight"]
        return qlib.conv2d_nhwc_tf32(x, w, self.w[name + ".bias"], 9, w.shape[0], res=res)

    def _conv1(self, x, name, res=None):
        w = self.w[name + ".weight"]
        return qlib.conv2d_nhwc_tf32(x, w, self.w[name + ".bias"], 1, w.shape[0], res=res)

    def _resblock(self, x, p):
        """QwenImageResidualBlock.forward (:242-286)."""
        h = self._conv1(x, p + ".conv_shortcut") if (p + ".conv_shortcut.weight") in self.w else x
        y = qlib.vae_rms_act(x, self.w[p + ".norm1.gamma"], True)
        t = self._conv3(y, p + ".conv1")
        y = qlib.vae_rms_act(t, self.w[p + ".norm2.gamma"], True, out=y if y.shape == t.shape else None)
        return self._conv3(y, p + ".conv2", res=h)

    def _attention(self, x, p):
        """QwenImageAttentionBlock.forward (:303-331): one head of width C over the h*w positions of each image."""
        N, H, W, Cc = x.shape
        P = H * W
        y = qlib.vae_rms_act(x, self.w[p + ".norm.gamma"], False)
        qkv = self._conv1(y, p + ".to_qkv")  # [N, H, W, 3C]
        att = torch.empty((N, H, W, Cc), dtype=torch.float32, device=x.device)
        band = max(8, min(H, (SCORES_BYTES_MAX // (4 * P * W)) // 8 * 8))  # query rows (of the image) per score band
        scores = torch.empty((min(band, H) * W, P), dtype=torch.float32, device=x.device)
        for n in range(N):
            k = qkv[n, :, :, Cc:2 * Cc]
            vt = qlib.vae_transpose(qkv[n].reshape(P, 3 * Cc)[:, 2 * Cc:])  # [C, P]
            for y0 in range(0, H, band):
                hb = min(band, H - y0)
                if hb < 8:  # the GEMM tile is a 16 x 8 pixel patch: extend the last band upwards
                    y0, hb = H - 8, 8
                q = qkv[n:n + 1, y0:y0 + hb, :, :Cc]
                s = scores[:hb * W].view(1, hb, W, P)
                qlib.conv2d_nhwc_tf32(q, _rows(k, P), None, 1, P, out=s, cin=Cc)
                qlib.vae_softmax_rows(s.view(hb * W, P), Cc ** -0.5)
                qlib.conv2d_nhwc_tf32(s, vt, None, 1, Cc, out=att[n:n + 1, y0:y0 + hb])
        return self._conv1(att, p + ".proj", res=x)

    # -- the reference surface --------------------------------------------------------------------------------------------
    def _features(self, z: torch.Tensor) -> torch.Tensor:
        """Everything up to and including norm_out + SiLU: [B, 8h, 8w, 96] NHWC."""
        if z.dim() != 5 or z.shape[2] != 1:
            raise ValueError("B200VaeDecoder decodes single-frame latents [B, z_dim, 1, h, w] (the image pipelines)")
        if not z.is_cuda:
            raise RuntimeError("B200VaeDecoder needs CUDA tensors: there is no CPU path")
        if z.shape[3] < 8 or z.shape[4] < 16:
            raise ValueError("latent grid must be at least 8 x 16 (images of 64 x 128 pixels)")
        x = qlib.vae_post_quant(z[:, :, 0].to(torch.float32).contiguous(), self.w["post_quant_conv.weight"],
                                self.w["post_quant_conv.bias"])
        x = self._conv3(x, "decoder.conv_in")
        x = self._resblock(x, "decoder.mid_block.resnets.0")
        x = self._attention(x, "decoder.mid_block.attentions.0")
        x = self._resblock(x, "decoder.mid_block.resnets.1")
        for i in range(self.num_up_blocks):
            for r in range(self.num_res):
                x = self._resblock(x, f"decoder.up_blocks.{i}.resnets.{r}")
            up = f"decoder.up_blocks.{i}.upsamplers.0.resample.1"
            if (up + ".weight") in self.w:
                x = self._conv3(qlib.vae_upsample2x(x), up)
        return qlib.vae_rms_act(x, self.w["decoder.norm_out.gamma"], True, out=x)

    # -- the reference surface --------------------------------------------------------------------------------------------
    @torch.no_grad()
    def decode(self, z: torch.Tensor, return_dict: bool = True):
        """z [B, z_dim, 1, h, w] (de-normalised latents) -> sample [B, 3, 1, 8h, 8w] fp32 in [-1, 1]."""
        y = self._features(z)
        img = qlib.vae_conv_out(y, self.w["decoder.conv_out.weight"], self.w["decoder.conv_out.bias"]).unsqueeze(2)
        return SimpleNamespace(sample=img) if return_dict else (img,)

    @torch.no_grad()
    def decode_to_uint8(self, z: torch.Tensor) -> torch.Tensor:
        """decode + the reference's post-process arithmetic (pipeline_qwen_image.py:40-60) in the last kernel:
        [B, 8h, 8w, 3] uint8 on the device, ready for `PIL.Image.fromarray` after one small device -> host copy."""
        y = self._features(z)
        return qlib.vae_conv_out(y, self.w["decoder.conv_out.weight"], self.w["decoder.conv_out.bias"], uint8=True)


class B200AutoencoderKLQwenImage(B200VaeDecoder):
    """Decode AND encode: adds `encode(x).latent_dist.mode()` — what the edit pipelines call on their condition image
    (pipeline_qwen_image_edit.py:458-480 -> AutoencoderKLQwenImage._encode, autoencoder_kl_qwenimage.py:793-812) — on the
    same kernels; the three resamplers are 3x3 stride-2 convolutions whose gather is the TMA descriptor's element stride."""
    PREFIXES = ("decoder.", "post_quant_conv.", "encoder.", "quant_conv.")

    def __init__(self, state_dict: dict, **kw):
        super().__init__(state_dict, **kw)
        if "encoder.conv_in.weight" not in self.w:
            raise ValueError("state dict holds no `encoder.*` tensors")
        self.num_down = 1 + max(int(k.split(".")[2]) for k in self.w if k.startswith("encoder.down_blocks."))

    @torch.no_grad()
    def _encode(self, x: torch.Tensor) -> torch.Tensor:
        """x [B, 3, 1, H, W] (or [B, 3, H, W]) in [-1, 1] -> posterior parameters [B, 2 z_dim, 1, H/8, W/8] fp32."""
        if x.dim() == 5:
            if x.shape[2] != 1:
                raise ValueError("single-frame images only")
            x = x[:, :, 0]
        if not x.is_cuda:
            raise RuntimeError("B200AutoencoderKLQwenImage needs CUDA tensors: there is no CPU path")
        B, C, H, W = x.shape
        if H % 8 or W % 8 or H < 64 or W < 128:
            raise ValueError("image height / width must be multiples of 8, at least 64 x 128")
        h = qlib.vae_image_to_nhwc(x.to(torch.float32).contiguous())
        h = self._conv3(h, "encoder.conv_in")
        for i in range(self.num_down):
            p = f"encoder.down_blocks.{i}"
            if (p + ".norm1.gamma") in self.w:
                h = self._resblock(h, p)
            else:  # downsample2d / downsample3d on the first frame (:190-199): zero-pad right / bottom, 3x3 stride 2
                w = self.w[p + ".resample.1.weight"]
                h = qlib.conv2d_down2_nhwc_tf32(h, w, self.w[p + ".resample.1.bias"], w.shape[0])
        h = self._resblock(h, "encoder.mid_block.resnets.0")
        h = self._attention(h, "encoder.mid_block.attentions.0")
        h = self._resblock(h, "encoder.mid_block.resnets.1")
        h = qlib.vae_rms_act(h, self.w["encoder.norm_out.gamma"], True, out=h)
        h = self._conv3(h, "encoder.conv_out")
        h = self._conv1(h, "quant_conv")                      # [B, H/8, W/8, 2 z]
        return h.permute(0, 3, 1, 2).unsqueeze(2).contiguous()  # the reference's NCHW(T) layout: 32 floats per latent pixel

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        dist = _DiagonalGaussian(self._encode(x))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)


class _DiagonalGaussian:
    """diffusers' DiagonalGaussianDistribution as far as the pipelines use it: `mode()` (sample_mode="argmax") and `sample`."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None else generator.device,
                            dtype=self.mean.dtype).to(self.mean.device)
        return self.mean + torch.exp(0.5 * self.logvar) * noise


def _rows(k: torch.Tensor, P: int) -> torch.Tensor:
    """[H, W, C] channel slice of the qkv buffer -> the [P, C] weight-side view (row stride = the buffer's pixel stride)."""
    return k.as_strided((P, k.shape[-1]), (k.stride(1), 1), k.storage_offset())
