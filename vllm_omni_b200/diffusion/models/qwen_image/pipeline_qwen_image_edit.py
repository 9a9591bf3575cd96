"""QwenImageEditPipeline — the image-conditioned sibling of QwenImagePipeline (reference
vllm_omni/diffusion/models/qwen_image/pipeline_qwen_image_edit.py): same transformer, same denoise loop, but the packed
VAE latents of the condition image are appended to the noisy latents on the sequence axis in every forward (:600-602),
`img_shapes` carries two grids per sample (:753-758) so RoPE gives the condition image frame index 1, and only the noisy
rows of the prediction are kept (:617,632).

Scope.  The Qwen2.5-VL prompt encode is outside the native engine.  The condition image comes either
already encoded —
    req.extra["image_latents"]      [B or 1, S2, 64] bf16 (what `prepare_latents` returns as `image_latents`, :519-522)
    req.extra["image_latent_grid"]  (h2, w2) latent-patch grid with h2 * w2 == S2
— or as pixels, when the injected `vae` can encode (`B200AutoencoderKLQwenImage`: native tcgen05 encoder):
    req.extra["image"]              [B or 1, 3, H, W] (or [.., 3, 1, H, W]) in [-1, 1], H and W multiples of 16: the output
                                    of the reference's pre-process (`VaeImageProcessor.preprocess`, :86-92)
which runs `_encode_vae_image` (:458-480: posterior mode, latent normalisation) and `_pack_latents` (:519-522).
"""
from __future__ import annotations

import torch

from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import (  # noqa: F401  (registry looks the func up here)
    QwenImagePipeline, get_qwen_image_post_process_func)
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest


class QwenImageEditPipeline(QwenImagePipeline):
    def diffuse(self, prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                image_latents, img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance,
                true_cfg_scale):
        """Reference signature (:574-589): `image_latents` is the sixth positional argument."""
        return self._denoise(prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                             img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale,
                             image_latents=image_latents)

    def _encode_vae_image(self, image: torch.Tensor):
        """Reference `_encode_vae_image` (:458-480) + `_pack_latents` (:519-522): image -> ([B, S2, 64] bf16, (h2, w2))."""
        if self.vae is None or not hasattr(self.vae, "encode"):
            raise ValueError("req.extra['image'] needs a VAE that can encode (B200AutoencoderKLQwenImage); pass image_latents instead")
        if image.dim() == 4:
            image = image.unsqueeze(2)
        lat = self.vae.encode(image.to(self.device)).latent_dist.mode()  # sample_mode="argmax" [B, 16, 1, h, w]
        z = self.vae.config.z_dim
        mean = torch.tensor(self.vae.config.latents_mean).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
        std = torch.tensor(self.vae.config.latents_std).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
        lat = (lat - mean) / std
        b, _, _, h, w = lat.shape
        if h % 2 or w % 2:
            raise ValueError("image height and width must be multiples of 16")
        return self._pack_latents(lat, b, z, h, w).to(torch.bfloat16), (h // 2, w // 2)

    def _condition_latents(self, req: OmniDiffusionRequest, batch: int, img_shapes):
        ex = req.extra or {}
        il = ex.get("image_latents")
        if il is None and ex.get("image") is not None:
            il, grid = self._encode_vae_image(ex["image"])
            ex = {**ex, "image_latent_grid": grid}
        if il is None:
            return None, img_shapes  # behaves as text-to-image, like the reference when `image` is None (:600-602)
        grid = ex.get("image_latent_grid")
        if grid is None or int(grid[0]) * int(grid[1]) != il.shape[1]:
            raise ValueError("req.extra['image_latent_grid'] = (h2, w2) with h2 * w2 == image_latents.shape[1] is required")
        if il.shape[0] != batch:
            if batch % il.shape[0]:
                raise ValueError(f"Cannot duplicate `image` of batch size {il.shape[0]} to {batch} text prompts.")  # (:512-515)
            il = torch.cat([il] * (batch // il.shape[0]), dim=0)
        shapes = [[img_shapes[0][0], (1, int(grid[0]), int(grid[1]))]] * batch
        return il, shapes


class QwenImageEditPlusPipeline(QwenImageEditPipeline):
    """Several condition images (reference pipeline_qwen_image_edit_plus.py:436-464,729-737): each image's packed latents
    follow the noisy latents on the sequence axis in the order given, `img_shapes` lists the output grid and then one grid
    per image, RoPE gives the k-th image the frame index k.  The request carries
        req.extra["image_latents"]       list of [B or 1, S_k, 64] tensors (or one tensor already concatenated on dim 1)
        req.extra["image_latent_grids"]  [(h_k, w_k), ...] with h_k * w_k == S_k
    A single image behaves exactly like QwenImageEditPipeline."""

    def _condition_latents(self, req: OmniDiffusionRequest, batch: int, img_shapes):
        ex = req.extra or {}
        il, grids = ex.get("image_latents"), ex.get("image_latent_grids")
        if il is None and isinstance(ex.get("image"), (list, tuple)):  # several condition images as pixels: encode each (:436-464)
            enc = [self._encode_vae_image(im) for im in ex["image"]]
            il, grids = [e[0] for e in enc], [e[1] for e in enc]
        if il is None and ex.get("image") is None:
            return None, img_shapes
        if grids is None:
            return super()._condition_latents(req, batch, img_shapes)
        parts = list(il) if isinstance(il, (list, tuple)) else list(torch.split(il, [int(h) * int(w) for h, w in grids], dim=1))
        if len(parts) != len(grids) or any(p.shape[1] != int(h) * int(w) for p, (h, w) in zip(parts, grids)):
            raise ValueError("image_latent_grids must list one (h, w) per condition image with h * w == its token count")
        rep = []
        for p in parts:
            if p.shape[0] != batch:
                if batch % p.shape[0]:
                    raise ValueError(f"Cannot duplicate `image` of batch size {p.shape[0]} to {batch} text prompts.")
                p = torch.cat([p] * (batch // p.shape[0]), dim=0)
            rep.append(p)
        shapes = [[img_shapes[0][0]] + [(1, int(h), int(w)) for h, w in grids]] * batch
        return torch.cat(rep, dim=1), shapes


def get_qwen_image_edit_post_process_func(od_config):
    return get_qwen_image_post_process_func(od_config)


def get_qwen_image_edit_plus_post_process_func(od_config):
    return get_qwen_image_post_process_func(od_config)
