"""B200-native Qwen-Image denoise pipeline — host-side mirror of the reference
`QwenImagePipeline` (vllm_omni/diffusion/models/qwen_image/pipeline_qwen_image.py:235-754)
for the part SURVEY.md §8 puts on the hot path: `prepare_latents`, `prepare_timesteps`,
the 50-step `diffuse()` loop with true-CFG + norm rescale + flow-match Euler step.

Prompt encoding (Qwen2.5-VL) and VAE decode run once per image, are unchanged reference
code and out of scope (§8f N1): this class takes pre-computed `prompt_embeds` and returns
latents (`output_type="latent"`) unless a text encoder / VAE object is injected.

Every arithmetic op runs in the C-ABI sm_100a library; the loop issues no host<->device
synchronisation (timesteps live on the device, sigmas are host floats).
"""
from __future__ import annotations

import logging
import math
from collections.abc import Iterable
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch
import torch.nn as nn

from vllm_omni_b200 import lib as qlib
from vllm_omni_b200.diffusion.data import DiffusionOutput, OmniDiffusionConfig
from vllm_omni_b200.diffusion.distributed import parallel_state as _ps
from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel
from vllm_omni_b200.diffusion.request import OmniDiffusionRequest

logger = logging.getLogger(__name__)


def calculate_shift(image_seq_len, base_seq_len: int = 256, max_seq_len: int = 4096, base_shift: float = 0.5,
                    max_shift: float = 1.15):
    """reference pipeline_qwen_image.py:63-73"""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


class FlowMatchEulerDiscreteScheduler:
    """Restatement of the diffusers scheduler the reference pipeline drives
    (pipeline_qwen_image.py:18-20,116-131,545,585): dynamic exponential time shift, terminal
    stretch, Euler step.  The config defaults are the Qwen-Image scheduler_config.json values
    recalled in SURVEY.md §8c (assumptions; callers may pass explicit sigmas)."""

    order = 1

    def __init__(self, num_train_timesteps: int = 1000, shift: float = 1.0, use_dynamic_shifting: bool = True,
                 base_shift: float = 0.5, max_shift: float = 0.9, base_image_seq_len: int = 256,
                 max_image_seq_len: int = 8192, shift_terminal: float | None = 0.02,
                 time_shift_type: str = "exponential"):
        self.config = dict(num_train_timesteps=num_train_timesteps, shift=shift, use_dynamic_shifting=use_dynamic_shifting,
                           base_shift=base_shift, max_shift=max_shift, base_image_seq_len=base_image_seq_len,
                           max_image_seq_len=max_image_seq_len, shift_terminal=shift_terminal,
                           time_shift_type=time_shift_type)
        self.sigmas: torch.Tensor | None = None      # fp32 [N+1] (host)
        self.timesteps: torch.Tensor | None = None   # fp32 [N]   (host)
        self._step_index: int | None = None
        self._begin_index: int | None = None

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def time_shift(self, mu: float, sigma: float, t: np.ndarray):
        if self.config["time_shift_type"] == "exponential":
            return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)
        return mu / (mu + (1 / t - 1) ** sigma)  # "linear"

    def stretch_shift_to_terminal(self, t: np.ndarray) -> np.ndarray:
        one_minus_z = 1 - t
        scale_factor = one_minus_z[-1] / (1 - self.config["shift_terminal"])
        return 1 - (one_minus_z / scale_factor)

    def set_timesteps(self, num_inference_steps: int | None = None, device=None, sigmas=None, mu: float | None = None):
        n_train = self.config["num_train_timesteps"]
        if sigmas is None:
            ts = np.linspace(n_train, 1, num_inference_steps)
            sigmas = ts / n_train
        sigmas = np.array(sigmas).astype(np.float32)
        if self.config["use_dynamic_shifting"]:
            if mu is None:
                raise ValueError("`mu` must be passed when `use_dynamic_shifting` is True")
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            s = self.config["shift"]
            sigmas = s * sigmas / (1 + (s - 1) * sigmas)
        if self.config["shift_terminal"]:
            sigmas = self.stretch_shift_to_terminal(sigmas)
        sig = torch.from_numpy(np.asarray(sigmas, dtype=np.float32))
        self.timesteps = sig * n_train
        self.sigmas = torch.cat([sig, torch.zeros(1)])
        self.num_inference_steps = len(sig)
        self._step_index = None

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = False):
        """x_{t-1} = bf16(float(x) + bf16(dt * v)) on the device (qimg_cfg_euler_step, no CFG)."""
        if self._step_index is None:
            self._step_index = self._begin_index or 0
        sigma, sigma_next = float(self.sigmas[self._step_index]), float(self.sigmas[self._step_index + 1])
        prev = sample.clone()
        qlib.cfg_euler_step(model_output.contiguous(), None, prev, 1.0, sigma, sigma_next)
        self._step_index += 1
        return (prev,)


@dataclass
class ComponentSource:
    """Same fields as DiffusersPipelineLoader.ComponentSource (reference diffusers_loader.py:40-60)."""

    model_or_path: str
    subfolder: str | None
    revision: str | None
    prefix: str = ""
    fall_back_to_pt: bool = True


class QwenImagePipeline(nn.Module):
    # prompt template of the reference pipeline (pipeline_qwen_image.py:283-285): the first 34 tokens of the encoded
    # sequence are the system preamble and are dropped from the embeddings
    prompt_template_encode = ("<|im_start|>system\nDescribe the image by detailing the color, shape, size, texture, quantity, text, "
                              "spatial relationships of the objects and background:<|im_end|>\n<|im_start|>user\n{}<|im_end|>\n"
                              "<|im_start|>assistant\n")
    prompt_template_encode_start_idx = 34
    tokenizer_max_length = 1024

    def __init__(self, *, od_config: OmniDiffusionConfig, prefix: str = "", text_encoder=None, vae=None, tokenizer=None,
                 transformer_kwargs: dict | None = None):
        super().__init__()
        self.od_config = od_config
        self.tokenizer = tokenizer
        self.weights_sources = [ComponentSource(model_or_path=od_config.model, subfolder="transformer", revision=None,
                                                prefix="transformer.", fall_back_to_pt=True)]
        self.scheduler = FlowMatchEulerDiscreteScheduler()
        self.text_encoder = text_encoder
        self.vae = vae
        self.transformer = QwenImageTransformer2DModel(od_config=od_config, **(transformer_kwargs or {}))
        self.vae_scale_factor = 8
        self.default_sample_size = 128
        self._interrupt = False
        self._attention_kwargs = None
        self._current_timestep = None
        self._num_timesteps = 0
        self._use_graph = False
        self._graphs: dict = {}

    @property
    def device(self):
        return self.transformer.img_in.weight.device

    @property
    def interrupt(self):
        return self._interrupt

    @property
    def attention_kwargs(self):
        return self._attention_kwargs

    def load_weights(self, weights: Iterable[tuple[str, torch.Tensor]]) -> set[str]:
        """(name, tensor) stream with the `transformer.` prefix of weights_sources (reference :752-754)."""
        def strip(ws):
            for n, t in ws:
                if n.startswith("transformer."):
                    yield n[len("transformer."):], t
        return {"transformer." + n for n in self.transformer.load_weights(strip(weights))}

    # ---- prompt encoding glue: reference :348-434 (the encoder / tokenizer are the transformers objects the reference
    # itself loads, `Qwen2_5_VLForConditionalGeneration` / `Qwen2Tokenizer` :264-272 — injected, not re-implemented) ----
    @staticmethod
    def _extract_masked_hidden(hidden_states: torch.Tensor, mask: torch.Tensor):
        keep = mask.bool()
        return torch.split(hidden_states[keep], keep.sum(dim=1).tolist(), dim=0)

    def _get_qwen_prompt_embeds(self, prompt, dtype: torch.dtype | None = None):
        """Chat-template the prompts, run the text encoder, take the last hidden state, drop each sample's padding and the
        34-token preamble, right-pad to the longest remainder -> (embeds [B, T, joint], mask [B, T] of ones / zeros)."""
        if self.text_encoder is None or self.tokenizer is None:
            raise ValueError("a text encoder and a tokenizer are required to encode prompts (inject them, or pass prompt_embeds)")
        dtype = dtype or self.text_encoder.dtype
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        drop = self.prompt_template_encode_start_idx
        tok = self.tokenizer([self.prompt_template_encode.format(p) for p in prompts], max_length=self.tokenizer_max_length + drop,
                             padding=True, truncation=True, return_tensors="pt").to(self.device)
        hidden = self.text_encoder(input_ids=tok.input_ids, attention_mask=tok.attention_mask,
                                   output_hidden_states=True).hidden_states[-1]
        rows = [h[drop:] for h in self._extract_masked_hidden(hidden, tok.attention_mask)]
        T = max(r.size(0) for r in rows)
        embeds = torch.stack([torch.cat([r, r.new_zeros(T - r.size(0), r.size(1))]) for r in rows])
        mask = torch.stack([torch.cat([torch.ones(r.size(0), dtype=torch.long, device=r.device),
                                       torch.zeros(T - r.size(0), dtype=torch.long, device=r.device)]) for r in rows])
        return embeds.to(dtype=dtype), mask

    def encode_prompt(self, prompt, num_images_per_prompt: int = 1, prompt_embeds: torch.Tensor | None = None,
                      prompt_embeds_mask: torch.Tensor | None = None, max_sequence_length: int = 1024):
        """reference :398-434: encode (unless embeddings are given), truncate, repeat per output image (prompt-major)."""
        prompts = [prompt] if isinstance(prompt, str) else prompt
        b = len(prompts) if prompt_embeds is None else prompt_embeds.shape[0]
        if prompt_embeds is None:
            prompt_embeds, prompt_embeds_mask = self._get_qwen_prompt_embeds(prompts)
        prompt_embeds, prompt_embeds_mask = prompt_embeds[:, :max_sequence_length], prompt_embeds_mask[:, :max_sequence_length]
        t = prompt_embeds.shape[1]
        prompt_embeds = prompt_embeds.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, t, -1)
        prompt_embeds_mask = prompt_embeds_mask.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, t)
        return prompt_embeds, prompt_embeds_mask

    # ---- reference :435-488 ---------------------------------------------------------------------
    @staticmethod
    def _pack_latents(latents, batch_size, num_channels_latents, height, width):
        latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
        latents = latents.permute(0, 2, 4, 1, 3, 5)
        return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)

    @staticmethod
    def _unpack_latents(latents, height, width, vae_scale_factor):
        batch_size, num_patches, channels = latents.shape
        height = 2 * (int(height) // (vae_scale_factor * 2))
        width = 2 * (int(width) // (vae_scale_factor * 2))
        latents = latents.view(batch_size, height // 2, width // 2, channels // 4, 2, 2)
        latents = latents.permute(0, 3, 1, 4, 2, 5)
        return latents.reshape(batch_size, channels // (2 * 2), 1, height, width)

    def prepare_latents(self, batch_size, num_channels_latents, height, width, dtype, device, generator, latents=None):
        height = 2 * (int(height) // (self.vae_scale_factor * 2))
        width = 2 * (int(width) // (self.vae_scale_factor * 2))
        if latents is not None:
            return latents.to(device=device, dtype=dtype)
        shape = (batch_size, 1, num_channels_latents, height, width)
        gdev = generator.device if isinstance(generator, torch.Generator) else "cpu"
        noise = torch.randn(shape, generator=generator if isinstance(generator, torch.Generator) else None,
                            device=gdev, dtype=dtype).to(device)
        return self._pack_latents(noise, batch_size, num_channels_latents, height, width)

    def prepare_timesteps(self, num_inference_steps, sigmas, image_seq_len):
        """reference :492-509"""
        sigmas = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps) if sigmas is None else sigmas
        cfg = self.scheduler.config
        mu = calculate_shift(image_seq_len, cfg.get("base_image_seq_len", 256), cfg.get("max_image_seq_len", 4096),
                             cfg.get("base_shift", 0.5), cfg.get("max_shift", 1.15))
        self.scheduler.set_timesteps(sigmas=sigmas, mu=mu)
        return self.scheduler.timesteps, len(self.scheduler.timesteps)

    # ---- the hot loop: reference :530-586 ---------------------------------------------------------
    def diffuse(self, prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale):
        """Reference signature (pipeline_qwen_image.py:530-544)."""
        return self._denoise(prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                             img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale)

    def _denoise(self, prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                 img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale,
                 image_latents=None):
        """`_denoise_once` under the attention guard: the default (fast) attention pipeline takes its exponentials against
        a running reference maximum and is exact unless a score jumps more than 2^100 above it inside one KV tile; such a
        launch raises a device flag (include/qimg_b200.h qimg_fmha_overflow).  The flag is read ONCE per denoise (4 bytes,
        the only synchronisation of the loop); if it is set — on any rank that shares this trajectory — the process
        switches to the exact pipeline for good and the trajectory is recomputed from the same initial latents, so the
        caller always receives exact-softmax results (reference semantics: attention/backends/sdpa.py:56-64)."""
        args = (prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents, img_shapes,
                txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale)
        mode = qlib.get_fmha_mode()
        if (mode & 7) != qlib.FMHA_FAST:
            return self._denoise_once(*args, image_latents=image_latents)
        qlib.fmha_overflow(reset=True)  # launches of other callers (layer-level plug-ins) must not leak into this decision
        out = self._denoise_once(*args, image_latents=image_latents)
        if not _ps.any_rank_in_model_group(qlib.fmha_overflow(reset=True), latents.device):
            return out
        logger.warning("attention scores left the fast pipeline's exact range (jump > 2^100 inside one KV tile); "
                       "switching to the exact attention pipeline and recomputing the denoise")
        qlib.set_fmha_mode(qlib.FMHA_EXACT | (mode & 8))
        hook = getattr(self.transformer, "_teacache", None)
        if hook is not None:
            hook.reset_state()
        return self._denoise_once(*args, image_latents=image_latents)

    # ---- CUDA-graph replay of the per-timestep launch sequence ------------------------------------
    def enable_cuda_graph(self, on: bool = True) -> None:
        """Capture ONE denoise timestep — the 60-block forward(s) (~550 launches each) + the fused CFG / Euler kernel —
        in a CUDA graph per (B, S_img, T, cfg) bucket and replay it for every timestep: the timestep embedding input and
        (sigma_i, sigma_{i+1}) live in device memory (`qimg_cfg_euler_step_dev`), so no launch has a per-step host
        argument.  Worth ~4 % at B=1 / TP where the forward is launch-bound; results are bit-identical to the eager loop.
        Not used with step caches (host-side decision), CFG parallel or the NCCL TP mode (host callbacks)."""
        self._use_graph = bool(on)
        if not on:
            self._graphs.clear()

    def _graph_eligible(self, do_true_cfg, image_latents) -> bool:
        tr = self.transformer
        return (self._use_graph and image_latents is None and tr._teacache is None
                and not (do_true_cfg and _ps.get_cfg_parallel_world_size() == 2)
                and not (tr.tp_size > 1 and tr.tp_comm == "nccl"))

    def _denoise_graph(self, prompt_embeds, negative_prompt_embeds, latents, img_shapes, txt_seq_lens, negative_txt_seq_lens,
                       timesteps, do_true_cfg, true_cfg_scale):
        dev = latents.device
        B, S1, _ = latents.shape
        T = prompt_embeds.shape[1]
        Tn = negative_prompt_embeds.shape[1] if do_true_cfg else 0
        key = (B, S1, T, Tn, bool(do_true_cfg), float(true_cfg_scale), repr(img_shapes[0]))
        st = self._graphs.get(key)
        self.transformer.do_true_cfg = do_true_cfg

        def step(s):
            pos = self.transformer(hidden_states=s["lat"], timestep=s["t"], encoder_hidden_states=s["pe"],
                                   encoder_hidden_states_mask=None, img_shapes=img_shapes, txt_seq_lens=txt_seq_lens,
                                   return_dict=False, uniform_timestep=True)[0]
            neg = None
            if do_true_cfg:
                neg = self.transformer(hidden_states=s["lat"], timestep=s["t"], encoder_hidden_states=s["ne"],
                                       encoder_hidden_states_mask=None, img_shapes=img_shapes,
                                       txt_seq_lens=negative_txt_seq_lens, return_dict=False, uniform_timestep=True)[0]
            qlib.cfg_euler_step_dev(pos, neg, s["lat"], float(true_cfg_scale), s["sig"])

        if st is None:
            st = {"lat": torch.empty_like(latents, dtype=torch.bfloat16), "pe": torch.empty_like(prompt_embeds, dtype=torch.bfloat16),
                  "ne": torch.empty_like(negative_prompt_embeds, dtype=torch.bfloat16) if do_true_cfg else None,
                  "t": torch.zeros(1, dtype=torch.bfloat16, device=dev), "sig": torch.zeros(2, dtype=torch.float32, device=dev)}
            st["lat"].copy_(latents)
            st["pe"].copy_(prompt_embeds)
            if do_true_cfg:
                st["ne"].copy_(negative_prompt_embeds)
            step(st)  # eager warm-up: workspaces, RoPE tables, TMA descriptors, peer-memory registration
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step(st)
            st["graph"] = g
            self._graphs[key] = st
        st["lat"].copy_(latents)
        st["pe"].copy_(prompt_embeds)
        if do_true_cfg:
            st["ne"].copy_(negative_prompt_embeds)
        sig = self.scheduler.sigmas
        sig2 = torch.stack([sig[:-1], sig[1:]], dim=1).to(torch.float32).to(dev)      # [N, 2]
        t_dev = (timesteps.to(torch.bfloat16) / 1000).to(dev).contiguous()            # same two bf16 roundings as the eager loop
        for i in range(len(timesteps)):
            if self.interrupt:
                continue
            self._current_timestep = timesteps[i]
            st["t"].copy_(t_dev[i:i + 1])
            st["sig"].copy_(sig2[i])
            st["graph"].replay()
        return st["lat"].clone()

    def _denoise_once(self, prompt_embeds, prompt_embeds_mask, negative_prompt_embeds, negative_prompt_embeds_mask, latents,
                      img_shapes, txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, guidance, true_cfg_scale,
                      image_latents=None):
        """The denoise loop behind `diffuse`.  `image_latents` [B,S2,64] (image-edit pipelines, reference pipeline_qwen_image_edit.py:600-602,617): the
        condition latents follow the noisy ones on the sequence axis in every forward, `img_shapes` lists both grids and
        only the first S1 rows of the prediction feed the CFG / scheduler step."""
        self.scheduler.set_begin_index(0)
        dev = latents.device
        if self._graph_eligible(do_true_cfg, image_latents):
            return self._denoise_graph(prompt_embeds.to(dev, torch.bfloat16), negative_prompt_embeds.to(dev, torch.bfloat16)
                                       if do_true_cfg else None, latents.to(dev, torch.bfloat16).contiguous(), img_shapes,
                                       txt_seq_lens, negative_txt_seq_lens, timesteps, do_true_cfg, true_cfg_scale)
        latents = latents.to(torch.bfloat16).contiguous().clone()
        s1 = latents.shape[1]
        model_in = latents
        if image_latents is not None:
            # one [B, S1+S2, 64] buffer: the condition rows are written once, the noisy rows refreshed per step
            model_in = torch.cat([latents, image_latents.to(dev, torch.bfloat16)], dim=1).contiguous()
        sig = self.scheduler.sigmas  # host fp32 [N+1]
        # `timestep = t.expand(B).to(latents.dtype)` then `/ 1000` (reference :552,558): same two bf16 roundings,
        # computed once for all steps; one row per step, shared by the whole batch.
        t_dev = (timesteps.to(torch.bfloat16) / 1000).to(dev).contiguous()
        self.transformer.do_true_cfg = do_true_cfg
        # CFG parallel: rank 0 of the CFG group runs the positive branch, rank 1 the negative one
        cfg_par = do_true_cfg and _ps.get_cfg_parallel_world_size() == 2
        if cfg_par:
            self.transformer.do_true_cfg = False  # one branch per rank: the step cache keeps a single state here
        cfg_rank = _ps.get_cfg_parallel_rank() if cfg_par else 0
        if cfg_par and cfg_rank == 1:
            prompt_embeds, prompt_embeds_mask, txt_seq_lens = negative_prompt_embeds, negative_prompt_embeds_mask, negative_txt_seq_lens
        for i in range(len(timesteps)):
            if self.interrupt:
                continue
            self._current_timestep = timesteps[i]
            if image_latents is not None and i > 0:
                model_in[:, :s1].copy_(latents)
            noise_pred = self.transformer(
                hidden_states=model_in, timestep=t_dev[i:i + 1], guidance=guidance,
                encoder_hidden_states_mask=prompt_embeds_mask, encoder_hidden_states=prompt_embeds, img_shapes=img_shapes,
                txt_seq_lens=txt_seq_lens, attention_kwargs=self.attention_kwargs, return_dict=False, uniform_timestep=True)[0]
            neg_noise_pred = None
            if cfg_par:
                noise_pred, neg_noise_pred = _ps.cfg_all_gather(noise_pred)
            elif do_true_cfg:
                neg_noise_pred = self.transformer(
                    hidden_states=model_in, timestep=t_dev[i:i + 1], guidance=guidance,
                    encoder_hidden_states_mask=negative_prompt_embeds_mask, encoder_hidden_states=negative_prompt_embeds,
                    img_shapes=img_shapes, txt_seq_lens=negative_txt_seq_lens, attention_kwargs=self.attention_kwargs,
                    return_dict=False, uniform_timestep=True)[0]
            if image_latents is not None:  # keep the noisy rows only (:617,632)
                noise_pred = noise_pred[:, :s1].contiguous()
                if neg_noise_pred is not None:
                    neg_noise_pred = neg_noise_pred[:, :s1].contiguous()
            # fused: comb = neg + s (pos - neg); rescale by ||pos|| / ||comb||; x += (sigma_{i+1} - sigma_i) v
            qlib.cfg_euler_step(noise_pred, neg_noise_pred, latents, float(true_cfg_scale), float(sig[i]), float(sig[i + 1]))
        return latents

    def _condition_latents(self, req: OmniDiffusionRequest, batch: int, img_shapes):
        """Text-to-image: no condition image.  Overridden by the edit pipeline."""
        return None, img_shapes

    # ---- request entry point: reference :588-750 ----------------------------------------------------
    @torch.inference_mode()
    def forward(self, req: OmniDiffusionRequest, prompt=None, negative_prompt=None, true_cfg_scale: float = 4.0,
                height: int | None = None, width: int | None = None, num_inference_steps: int = 50,
                sigmas: list[float] | None = None, guidance_scale: float = 1.0, num_images_per_prompt: int = 1,
                generator=None, latents=None, prompt_embeds=None, prompt_embeds_mask=None, negative_prompt_embeds=None,
                negative_prompt_embeds_mask=None, output_type: str | None = "latent",
                attention_kwargs: dict[str, Any] | None = None, max_sequence_length: int = 512) -> DiffusionOutput:
        height = req.height or height or self.default_sample_size * self.vae_scale_factor
        width = req.width or width or self.default_sample_size * self.vae_scale_factor
        num_inference_steps = req.num_inference_steps or num_inference_steps
        generator = req.generator or generator
        if generator is None and req.seed is not None:
            generator = torch.Generator().manual_seed(req.seed)
        true_cfg_scale = req.true_cfg_scale or true_cfg_scale
        sigmas = req.sigmas if req.sigmas is not None else sigmas
        latents = req.latents if req.latents is not None else latents
        output_type = req.output_type or output_type
        if req.num_outputs_per_prompt and req.num_outputs_per_prompt > 0:
            num_images_per_prompt = req.num_outputs_per_prompt
        prompt_embeds = req.prompt_embeds if req.prompt_embeds is not None else prompt_embeds
        negative_prompt_embeds = req.negative_prompt_embeds if req.negative_prompt_embeds is not None else negative_prompt_embeds
        prompt_embeds_mask = req.prompt_attention_mask if req.prompt_attention_mask is not None else prompt_embeds_mask
        negative_prompt_embeds_mask = (req.negative_attention_mask if req.negative_attention_mask is not None
                                       else negative_prompt_embeds_mask)
        if height % (self.vae_scale_factor * 2) or width % (self.vae_scale_factor * 2):
            raise ValueError(f"height and width must be divisible by {self.vae_scale_factor * 2}")
        if prompt_embeds is None:  # in-process prompt encoding, as the reference does (:357-396); else the embeddings come
            # from the request (an upstream encoder stage feeds them through a connector)
            prompt_embeds, prompt_embeds_mask = self._get_qwen_prompt_embeds(req.prompt if req.prompt is not None else prompt)
            neg = req.negative_prompt if req.negative_prompt is not None else negative_prompt
            if neg is not None and true_cfg_scale > 1:
                negative_prompt_embeds, negative_prompt_embeds_mask = self._get_qwen_prompt_embeds(neg)
        dev = self.device

        def rep(e, m):
            e = e.to(dev, torch.bfloat16)
            if m is None:
                m = torch.ones(e.shape[:2], dtype=torch.long)
            b, s, _ = e.shape
            e = e.repeat(1, num_images_per_prompt, 1).view(b * num_images_per_prompt, s, -1)
            m = m.repeat(1, num_images_per_prompt).view(b * num_images_per_prompt, s)
            return e.contiguous(), m

        self._attention_kwargs = attention_kwargs or {}
        self._interrupt = False
        batch_size = prompt_embeds.shape[0]
        do_true_cfg = true_cfg_scale > 1 and negative_prompt_embeds is not None
        prompt_embeds, prompt_embeds_mask = rep(prompt_embeds, prompt_embeds_mask)
        if do_true_cfg:
            negative_prompt_embeds, negative_prompt_embeds_mask = rep(negative_prompt_embeds, negative_prompt_embeds_mask)
        latents = self.prepare_latents(batch_size * num_images_per_prompt, self.transformer.in_channels // 4, height, width,
                                       torch.bfloat16, dev, generator, latents)
        img_shapes = [[(1, height // self.vae_scale_factor // 2, width // self.vae_scale_factor // 2)]] * (
            batch_size * num_images_per_prompt)
        image_latents, img_shapes = self._condition_latents(req, batch_size * num_images_per_prompt, img_shapes)
        timesteps, num_inference_steps = self.prepare_timesteps(num_inference_steps, sigmas, latents.shape[1])
        self._num_timesteps = len(timesteps)
        txt_seq_lens = prompt_embeds_mask.sum(dim=1).tolist()
        negative_txt_seq_lens = negative_prompt_embeds_mask.sum(dim=1).tolist() if do_true_cfg else None
        latents = self._denoise(prompt_embeds, prompt_embeds_mask, negative_prompt_embeds if do_true_cfg else None,
                                negative_prompt_embeds_mask if do_true_cfg else None, latents, img_shapes, txt_seq_lens,
                                negative_txt_seq_lens, timesteps, do_true_cfg, None, true_cfg_scale,
                                image_latents=image_latents)
        self._current_timestep = None
        if output_type == "latent" or self.vae is None:
            return DiffusionOutput(output=latents)
        return DiffusionOutput(output=self.decode_latents(self.vae, latents, height, width, self.vae_scale_factor,
                                                          uint8=getattr(self, "uint8_output", False)))

    @staticmethod
    def decode_latents(vae, latents, height, width, vae_scale_factor: int = 8, uint8: bool = False):
        """The reference's post-step (:736-747): unpack, cast to the VAE's dtype, undo the latent normalisation, decode, keep
        frame 0.  `vae` is anything with the AutoencoderKLQwenImage decode surface — the native `B200VaeDecoder`
        (vae_decoder.py: tcgen05 TF32 convolutions) or a torch module."""
        lat = QwenImagePipeline._unpack_latents(latents, height, width, vae_scale_factor).to(vae.dtype)
        z = vae.config.z_dim
        mean = torch.tensor(vae.config.latents_mean).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
        std = 1.0 / torch.tensor(vae.config.latents_std).view(1, z, 1, 1, 1).to(lat.device, lat.dtype)
        if uint8 and hasattr(vae, "decode_to_uint8"):  # native decoder: post-process arithmetic fused into its last kernel
            return vae.decode_to_uint8(lat / std + mean)
        return vae.decode(lat / std + mean, return_dict=False)[0][:, :, 0]


def get_qwen_image_post_process_func(od_config: OmniDiffusionConfig):
    """Looked up by name by the registry (reference registry.py:97-139).  The reference's function is
    `VaeImageProcessor.postprocess` (pipeline :40-60, diffusers): (x / 2 + 0.5).clamp(0, 1) -> NHWC -> (x * 255).round()
    -> uint8 -> PIL.  Here: latents ([B, S, 64], no VAE in the worker) pass through; a decoded image [B, 3, H, W] in
    [-1, 1] gets that arithmetic; a uint8 [B, H, W, 3] tensor (B200VaeDecoder.decode_to_uint8: the arithmetic already done on
    the device) is only copied to the host."""
    def post_process_func(images: torch.Tensor):
        if not torch.is_tensor(images) or images.dim() != 4:
            return images
        if images.dtype != torch.uint8:
            images = ((images.float() / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8)
        from PIL import Image
        return [Image.fromarray(a) for a in images.cpu().numpy()]
    return post_process_func
