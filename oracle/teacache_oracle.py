"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's TeaCache step cache for the Qwen-Image transformer
(vllm_omni/diffusion/cache/teacache/: hook.py:80-217, extractors.py:184-246, config.py:9-29, state.py), on top of the
oracle's staged model forward.  Pinned against the unmodified reference hook by oracle/make_golden_teacache.py
(fixture tests/golden/teacache_tiny.pt)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from oracle import qwen_image_oracle as O

# config.py:19-29 ("Qwen-Image transformer coefficients from ComfyUI-TeaCache")
QWEN_IMAGE_COEFFICIENTS = [-4.50000000e02, 2.80000000e02, -4.50000000e01, 3.20000000e00, -2.00000000e-02]


class TeaCacheState:
    """state.py:17-38."""

    def __init__(self):
        self.cnt = 0
        self.accumulated_rel_l1_distance = 0.0
        self.previous_modulated_input = None
        self.previous_residual = None
        self.previous_residual_encoder = None


class TeaCacheOracle:
    def __init__(self, w: dict, dims: O.DiTDims, rel_l1_thresh: float = 0.2, coefficients=None):
        self.w, self.dims, self.thresh = w, dims, rel_l1_thresh
        self.rescale = np.poly1d(QWEN_IMAGE_COEFFICIENTS if coefficients is None else coefficients)  # hook.py:57
        self.states = {"positive": TeaCacheState(), "negative": TeaCacheState()}
        self.forward_cnt = 0
        self.decisions: list[tuple[str, bool, float]] = []  # (branch, computed?, rel distance)

    def reset(self):
        self.states = {"positive": TeaCacheState(), "negative": TeaCacheState()}
        self.forward_cnt = 0
        self.decisions = []

    def should_compute(self, st: TeaCacheState, mod: torch.Tensor):
        """hook.py:170-217."""
        if st.cnt == 0:
            st.accumulated_rel_l1_distance = 0.0
            return True, float("nan")
        if st.previous_modulated_input is None:
            return True, float("nan")
        rel = ((mod - st.previous_modulated_input).abs().mean()
               / (st.previous_modulated_input.abs().mean() + 1e-8)).cpu().item()
        st.accumulated_rel_l1_distance += abs(float(self.rescale(rel)))
        if st.accumulated_rel_l1_distance < self.thresh:
            return False, rel
        st.accumulated_rel_l1_distance = 0.0
        return True, rel

    def forward(self, hidden_states, encoder_hidden_states, timestep, img_shape, do_true_cfg: bool = False):
        """hook.new_forward (:80-165) with the Qwen extractor (extractors.py:184-246)."""
        w, dims = self.w, self.dims
        img, txt, temb, rope = O.model_pre(w, dims, hidden_states, encoder_hidden_states, timestep, img_shape)
        p = "transformer_blocks.0."
        img_mod1 = F.linear(F.silu(temb), w[p + "img_mod.1.weight"], w[p + "img_mod.1.bias"]).chunk(2, dim=-1)[0]
        mod, _ = O.ada_layer_norm(img, img_mod1, dims.eps)  # extractors.py:206-209
        branch = "negative" if (do_true_cfg and self.forward_cnt % 2 == 1) else "positive"  # hook.py:115-119
        st = self.states[branch]
        compute, rel = self.should_compute(st, mod)
        if not compute and st.previous_residual is not None:
            img = img + st.previous_residual  # :131
        else:
            ori_img, ori_txt = img.clone(), txt.clone()
            img, txt = O.model_blocks(w, dims, img, txt, temb, rope)
            st.previous_residual = img - ori_img  # :152
            st.previous_residual_encoder = txt - ori_txt
            compute = True
        st.previous_modulated_input = mod
        st.cnt += 1
        self.forward_cnt += 1
        self.decisions.append((branch, compute, rel))
        return O.model_post(w, dims, img, temb)


def diffuse(tc: TeaCacheOracle, latents, prompt_embeds, neg_prompt_embeds, sigmas, img_shape, true_cfg_scale: float = 4.0):
    """QwenImagePipeline.diffuse (pipeline_qwen_image.py:530-586) with the hooked transformer."""
    sig = torch.from_numpy(np.asarray(sigmas, dtype=np.float32))
    timesteps = sig[:-1] * 1000.0
    do_cfg = neg_prompt_embeds is not None
    for i, t in enumerate(timesteps):
        timestep = t.expand(latents.shape[0]).to(dtype=latents.dtype)
        noise = tc.forward(latents, prompt_embeds, timestep / 1000, img_shape, do_cfg)
        if do_cfg:
            neg = tc.forward(latents, neg_prompt_embeds, timestep / 1000, img_shape, do_cfg)
            noise = O.cfg_combine(noise, neg, true_cfg_scale)
        latents = O.euler_step(noise, latents, sig[i], sig[i + 1])
    return latents
