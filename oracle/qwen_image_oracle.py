"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch, dtype-generic) of the
reference's Qwen-Image DiT denoising hot path.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / `--impl reference` leg may import this module, and only
as the checker / CPU baseline — never as the thing shipped.  The product
(vllm_omni_b200) must not import anything under oracle/.

Every function cites the reference file:line it follows (paths relative to the
vllm-omni tree, commit be81443).  The same code runs in two modes:
  * fp32 tensors  -> the fp32 oracle ("ground truth");
  * bf16 tensors  -> reproduces the reference's own op-by-op bf16 rounding, because
    it issues the same torch ops in the same order as the reference modules do.

Pinning: the reference's tests hold NO golden vectors for this path.
The restatement is pinned instead against outputs of the reference itself, run in the
build container through oracle/ref_shim.py: tests/golden/*.pt were produced by
oracle/make_golden.py from the unmodified reference classes, and
tests/test_oracle_vs_golden.py checks this file against them (bit-exact in bf16 on
the same torch build).  The diffusers-owned pieces (FeedForward, Timesteps,
TimestepEmbedding, AdaLayerNormContinuous, FlowMatchEulerDiscreteScheduler) are absent
from the image; they are restated from the published algorithm (diffusers>=0.36.0,
pyproject.toml:35) and are "parity unpinned" at that boundary.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# dimensions
# --------------------------------------------------------------------------------------
@dataclass
class DiTDims:
    """Constructor defaults of QwenImageTransformer2DModel (qwen_image_transformer.py:635-650)."""

    num_layers: int = 60
    num_heads: int = 24
    head_dim: int = 128
    in_channels: int = 64
    out_channels: int = 16
    patch_size: int = 2
    joint_dim: int = 3584
    axes_dims_rope: tuple = (16, 56, 56)
    eps: float = 1e-6

    @property
    def dim(self) -> int:
        return self.num_heads * self.head_dim

    @property
    def proj_out_dim(self) -> int:
        return self.patch_size * self.patch_size * self.out_channels


# --------------------------------------------------------------------------------------
# small ops
# --------------------------------------------------------------------------------------
def timestep_sinusoid(timesteps: torch.Tensor, dim: int = 256, scale: float = 1000.0) -> torch.Tensor:
    """diffusers Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000)
    (qwen_image_transformer.py:44); restated in-tree at pipeline_qwen_image.py:135-184.
    Returns fp32 [N, dim] = [cos | sin]."""
    half = dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=timesteps.device)
    exponent = exponent / (half - 0.0)
    emb = torch.exp(exponent)
    emb = timesteps[:, None].float() * emb[None, :]
    emb = scale * emb
    emb = torch.cat([torch.sin(emb), torch.cos(emb)], dim=-1)
    emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)  # flip_sin_to_cos
    return emb


def time_text_embed(w: dict, timestep: torch.Tensor, dtype) -> torch.Tensor:
    """QwenTimestepProjEmbeddings.forward (qwen_image_transformer.py:50-62)."""
    proj = timestep_sinusoid(timestep).to(dtype)
    h = F.linear(proj, w["time_text_embed.timestep_embedder.linear_1.weight"],
                 w["time_text_embed.timestep_embedder.linear_1.bias"])
    h = F.silu(h)
    return F.linear(h, w["time_text_embed.timestep_embedder.linear_2.weight"],
                    w["time_text_embed.timestep_embedder.linear_2.bias"])


def rms_norm(x: torch.Tensor, weight: torch.Tensor, eps: float) -> torch.Tensor:
    """vLLM RMSNorm native path (installed vllm/ir/ops/layernorm.py:10-21) used at
    qwen_image_transformer.py:324-325,353-354,669."""
    orig = x.dtype
    xf = x.to(torch.float32)
    var = xf.pow(2).mean(dim=-1, keepdim=True)
    xf = xf * torch.rsqrt(var + eps)
    return (xf.to(weight.dtype) * weight).to(orig)


def ada_layer_norm(x: torch.Tensor, mod: torch.Tensor, eps: float, index: torch.Tensor = None):
    """AdaLayerNorm.forward_native (layers/adalayernorm.py:94-102); chunk order
    (shift, scale, gate) from :29.  x [B,S,D], mod [B,3D] -> (y [B,S,D], gate [B,1,D]).
    With `index` [B,S] (`preprocess`, :31-54): mod has 2B rows, token (b, s) takes row b when index == 0 and row B + b
    otherwise; the gate is then per token [B,S,D]."""
    shift, scale, gate = mod.chunk(3, dim=-1)
    if index is not None:
        B = shift.size(0) // 2
        sel = (index == 0).unsqueeze(-1)
        shift = torch.where(sel, shift[:B].unsqueeze(1), shift[B:].unsqueeze(1))
        scale = torch.where(sel, scale[:B].unsqueeze(1), scale[B:].unsqueeze(1))
        gate = torch.where(sel, gate[:B].unsqueeze(1), gate[B:].unsqueeze(1))
    else:
        shift, scale, gate = shift.unsqueeze(1), scale.unsqueeze(1), gate.unsqueeze(1)
    y = F.layer_norm(x, (x.shape[-1],), None, None, eps) * (1 + scale) + shift
    return y, gate


def rope_tables(frame, height: int = None, width: int = None, txt_len: int = None, axes=(16, 56, 56), theta: float = 10000.0):
    """QwenEmbedRope (qwen_image_transformer.py:179-285) with scale_rope=True: returns fp32
    (img_cos, img_sin [sum F*H*W, 64], txt_cos, txt_sin [T, 64]); cos/sin are the real/imag parts of
    torch.polar(1, angle) (:220).  `frame, height, width` describe ONE image; alternatively pass a list of
    (frame, height, width) grids as the first argument (image-edit pipelines: noisy latents followed by the condition
    image(s), QwenEmbedRope.forward :222-260): the idx-th grid takes its frame positions from idx (:267), the tables are
    concatenated along the sequence (:258) and the text positions start after the largest grid (:251-257)."""
    if isinstance(frame, (list, tuple)) and len(frame) and isinstance(frame[0], (list, tuple)):
        grids, txt_len = [tuple(g) for g in frame], (height if txt_len is None else txt_len)
    else:
        grids = [(frame, height, width)]
    def params(index, dim):
        freqs = torch.outer(index.float(), 1.0 / torch.pow(theta, torch.arange(0, dim, 2).to(torch.float32).div(dim)))
        return freqs  # angles; polar(1, a) = cos a + i sin a

    pos_index = torch.arange(4096)
    neg_index = torch.arange(4096).flip(0) * -1 - 1
    pos = [params(pos_index, d) for d in axes]
    neg = [params(neg_index, d) for d in axes]
    angs, max_vid_index = [], 0
    for idx, (frame, height, width) in enumerate(grids):
        # _compute_video_freqs (:262-285)
        f_frame = pos[0][idx: idx + frame].view(frame, 1, 1, -1).expand(frame, height, width, -1)
        f_h = torch.cat([neg[1][-(height - height // 2):], pos[1][: height // 2]], dim=0)
        f_h = f_h.view(1, height, 1, -1).expand(frame, height, width, -1)
        f_w = torch.cat([neg[2][-(width - width // 2):], pos[2][: width // 2]], dim=0)
        f_w = f_w.view(1, 1, width, -1).expand(frame, height, width, -1)
        angs.append(torch.cat([f_frame, f_h, f_w], dim=-1).reshape(frame * height * width, -1))
        max_vid_index = max(height // 2, width // 2, max_vid_index)  # (:251-254)
    ang = torch.cat(angs, dim=0)
    txt_ang = torch.cat(pos, dim=1)[max_vid_index: max_vid_index + txt_len]  # (:257)
    # torch.polar(ones, a): real = cos a, imag = sin a in fp32
    return torch.cos(ang), torch.sin(ang), torch.cos(txt_ang), torch.sin(txt_ang)


def apply_rope_interleaved(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """RotaryEmbedding(is_neox_style=False).forward_native -> apply_rotary_emb_torch(interleaved=True)
    (layers/rope.py:13-36,142-151).  x [B,S,H,hd], cos/sin [S,hd/2] already in x.dtype
    (qwen_image_transformer.py:403-406)."""
    cos2 = cos.repeat_interleave(2, dim=-1)[:, None, :]  # "... d -> ... 1 (d 2)"
    sin2 = sin.repeat_interleave(2, dim=-1)[:, None, :]
    x1, x2 = x[..., ::2], x[..., 1::2]
    rot = torch.stack((-x2, x1), dim=-1).flatten(-2)
    return x * cos2 + rot * sin2


def joint_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
    """Attention -> SDPAImpl.forward (attention/layer.py:54-70, backends/sdpa.py:46-66):
    q,k,v [B,S,H,hd] -> [B,S,H,hd]; non-causal, no mask."""
    q, k, v = (t.permute(0, 2, 1, 3) for t in (q, k, v))
    o = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=scale)
    return o.permute(0, 2, 1, 3)


def feed_forward(x, w: dict, prefix: str):
    """diffusers FeedForward(dim, dim_out=dim, activation_fn='gelu-approximate')
    (qwen_image_transformer.py:491,501): net.0.proj -> gelu(tanh) -> net.2."""
    h = F.gelu(F.linear(x, w[prefix + ".net.0.proj.weight"], w[prefix + ".net.0.proj.bias"]), approximate="tanh")
    return F.linear(h, w[prefix + ".net.2.weight"], w[prefix + ".net.2.bias"])


# --------------------------------------------------------------------------------------
# block / model
# --------------------------------------------------------------------------------------
def cross_attention(w: dict, p: str, dims: DiTDims, img, txt, rope):
    """QwenImageCrossAttention.forward (qwen_image_transformer.py:370-458).  `p` is the
    block prefix ('transformer_blocks.3.')."""
    H, hd = dims.num_heads, dims.head_dim
    T = txt.shape[1]
    img_cos, img_sin, txt_cos, txt_sin = rope
    qkv = F.linear(img, w[p + "attn.to_qkv.weight"], w[p + "attn.to_qkv.bias"])
    iq, ik, iv = (t.unflatten(-1, (H, -1)) for t in qkv.chunk(3, dim=-1))
    qkv = F.linear(txt, w[p + "attn.add_kv_proj.weight"], w[p + "attn.add_kv_proj.bias"])
    tq, tk, tv = (t.unflatten(-1, (H, -1)) for t in qkv.chunk(3, dim=-1))
    iq = rms_norm(iq, w[p + "attn.norm_q.weight"], dims.eps)
    ik = rms_norm(ik, w[p + "attn.norm_k.weight"], dims.eps)
    tq = rms_norm(tq, w[p + "attn.norm_added_q.weight"], dims.eps)
    tk = rms_norm(tk, w[p + "attn.norm_added_k.weight"], dims.eps)
    dt = iq.dtype
    iq = apply_rope_interleaved(iq, img_cos.to(dt), img_sin.to(dt))
    ik = apply_rope_interleaved(ik, img_cos.to(dt), img_sin.to(dt))
    tq = apply_rope_interleaved(tq, txt_cos.to(dt), txt_sin.to(dt))
    tk = apply_rope_interleaved(tk, txt_cos.to(dt), txt_sin.to(dt))
    jq = torch.cat([tq, iq], dim=1)
    jk = torch.cat([tk, ik], dim=1)
    jv = torch.cat([tv, iv], dim=1)
    o = joint_attention(jq, jk, jv, 1.0 / (hd ** 0.5)).flatten(2, 3).to(jq.dtype)
    txt_o, img_o = o[:, :T, :], o[:, T:, :]
    img_o = F.linear(img_o, w[p + "attn.to_out.0.weight"], w[p + "attn.to_out.0.bias"])
    txt_o = F.linear(txt_o, w[p + "attn.to_add_out.weight"], w[p + "attn.to_add_out.bias"])
    return img_o, txt_o


def block_forward(w: dict, p: str, dims: DiTDims, img, txt, temb, rope):
    """QwenImageTransformerBlock.forward (qwen_image_transformer.py:541-605), zero_cond_t=False."""
    img_mod = F.linear(F.silu(temb), w[p + "img_mod.1.weight"], w[p + "img_mod.1.bias"])
    txt_mod = F.linear(F.silu(temb), w[p + "txt_mod.1.weight"], w[p + "txt_mod.1.bias"])
    img_mod1, img_mod2 = img_mod.chunk(2, dim=-1)
    txt_mod1, txt_mod2 = txt_mod.chunk(2, dim=-1)
    img_m, img_g1 = ada_layer_norm(img, img_mod1, dims.eps)
    txt_m, txt_g1 = ada_layer_norm(txt, txt_mod1, dims.eps)
    img_a, txt_a = cross_attention(w, p, dims, img_m, txt_m, rope)
    img = img + img_g1 * img_a
    txt = txt + txt_g1 * txt_a
    img_m2, img_g2 = ada_layer_norm(img, img_mod2, dims.eps)
    img = img + img_g2 * feed_forward(img_m2, w, p + "img_mlp")
    txt_m2, txt_g2 = ada_layer_norm(txt, txt_mod2, dims.eps)
    txt = txt + txt_g2 * feed_forward(txt_m2, w, p + "txt_mlp")
    return txt, img


def model_pre(w: dict, dims: DiTDims, hidden_states, encoder_hidden_states, timestep, img_shape, txt_len=None):
    """Preprocessing part of QwenImageTransformer2DModel.forward (qwen_image_transformer.py:743-770; the same lines the
    TeaCache extractor runs, cache/teacache/extractors.py:187-204): -> (img, txt, temb, rope)."""
    dt = hidden_states.dtype
    img = F.linear(hidden_states, w["img_in.weight"], w["img_in.bias"])
    timestep = timestep.to(dt)
    txt = rms_norm(encoder_hidden_states, w["txt_norm.weight"], dims.eps)
    txt = F.linear(txt, w["txt_in.weight"], w["txt_in.bias"])
    temb = time_text_embed(w, timestep, dt)
    T = encoder_hidden_states.shape[1] if txt_len is None else txt_len
    if isinstance(img_shape[0], (list, tuple)):  # several grids per sample (edit pipelines)
        rope = rope_tables([tuple(g) for g in img_shape], T, axes=dims.axes_dims_rope)
    else:
        rope = rope_tables(*img_shape, T, axes=dims.axes_dims_rope)
    return img, txt, temb, rope


def model_blocks(w: dict, dims: DiTDims, img, txt, temb, rope, inter: list | None = None):
    """The transformer blocks (:783-792; extractor `run_transformer_blocks`, :214-227): -> (img, txt)."""
    for i in range(dims.num_layers):
        txt, img = block_forward(w, f"transformer_blocks.{i}.", dims, img, txt, temb, rope)
        if inter is not None:
            inter.append((txt, img))
    return img, txt


def model_post(w: dict, dims: DiTDims, img, temb):
    """AdaLayerNormContinuous (:686,797; scale FIRST, then shift) + proj_out (:798); extractor `postprocess` :232-238."""
    dt = img.dtype
    emb = F.linear(F.silu(temb).to(dt), w["norm_out.linear.weight"], w["norm_out.linear.bias"])
    scale, shift = torch.chunk(emb, 2, dim=1)
    img = F.layer_norm(img, (img.shape[-1],), None, None, dims.eps) * (1 + scale)[:, None, :] + shift[:, None, :]
    return F.linear(img, w["proj_out.weight"], w["proj_out.bias"])


def model_forward(w: dict, dims: DiTDims, hidden_states, encoder_hidden_states, timestep, img_shape, txt_len=None,
                  return_intermediates: bool = False):
    """QwenImageTransformer2DModel.forward (qwen_image_transformer.py:692-802), SP off,
    zero_cond_t off, guidance None.  hidden_states [B,S_img,64]; encoder_hidden_states
    [B,T,joint]; timestep [B] (already /1000); img_shape = (frame, h, w) latent-patch grid (or a list of grids)."""
    img, txt, temb, rope = model_pre(w, dims, hidden_states, encoder_hidden_states, timestep, img_shape, txt_len)
    inter = [] if return_intermediates else None
    img, txt = model_blocks(w, dims, img, txt, temb, rope, inter)
    out = model_post(w, dims, img, temb)
    if return_intermediates:
        return out, inter
    return out


# --------------------------------------------------------------------------------------
# denoise loop pieces
# --------------------------------------------------------------------------------------
def calculate_shift(image_seq_len, base_seq_len=256, max_seq_len=4096, base_shift=0.5, max_shift=1.15):
    """pipeline_qwen_image.py:63-73."""
    m = (max_shift - base_shift) / (max_seq_len - base_seq_len)
    b = base_shift - m * base_seq_len
    return image_seq_len * m + b


# Qwen-Image scheduler_config.json is NOT in the reference tree; these values are the
# ones recalled from the HF repo and are ASSUMPTIONS.  Parity
# never depends on them because benches/tests pass the sigma table explicitly.
QWEN_IMAGE_SCHED = dict(num_train_timesteps=1000, use_dynamic_shifting=True, base_shift=0.5, max_shift=0.9,
                        base_image_seq_len=256, max_image_seq_len=8192, shift_terminal=0.02,
                        time_shift_type="exponential")


def flow_match_sigmas(num_steps: int, image_seq_len: int, cfg: dict = QWEN_IMAGE_SCHED) -> np.ndarray:
    """prepare_timesteps (pipeline_qwen_image.py:492-509) + published
    FlowMatchEulerDiscreteScheduler.set_timesteps(sigmas=..., mu=...): exponential time shift,
    terminal stretch, float32, append 0.  Returns float32 [num_steps + 1]."""
    sigmas = np.linspace(1.0, 1 / num_steps, num_steps)
    mu = calculate_shift(image_seq_len, cfg["base_image_seq_len"], cfg["max_image_seq_len"], cfg["base_shift"],
                         cfg["max_shift"])
    sigmas = np.array(sigmas).astype(np.float32)
    sigmas = math.exp(mu) / (math.exp(mu) + (1 / sigmas - 1) ** 1.0)  # time_shift exponential
    if cfg.get("shift_terminal"):
        one_minus = 1 - sigmas
        scale_factor = one_minus[-1] / (1 - cfg["shift_terminal"])
        sigmas = 1 - (one_minus / scale_factor)
    sigmas = sigmas.astype(np.float32)
    return np.concatenate([sigmas, np.zeros(1, dtype=np.float32)])


def cfg_combine(pos: torch.Tensor, neg: torch.Tensor, true_cfg_scale: float) -> torch.Tensor:
    """pipeline_qwen_image.py:580-583."""
    comb = neg + true_cfg_scale * (pos - neg)
    cond_norm = torch.norm(pos, dim=-1, keepdim=True)
    noise_norm = torch.norm(comb, dim=-1, keepdim=True)
    return comb * (cond_norm / noise_norm)


def euler_step(noise_pred: torch.Tensor, latents: torch.Tensor, sigma: torch.Tensor, sigma_next: torch.Tensor):
    """FlowMatchEulerDiscreteScheduler.step (called at pipeline_qwen_image.py:585):
    sample upcast to fp32; dt = sigma_next - sigma is a 0-dim fp32 tensor, so under torch
    type promotion `dt * model_output` is computed in model_output.dtype (dt itself is rounded
    to bf16 first when the model runs in bf16); the sum is fp32 and is cast back to
    model_output.dtype."""
    sample = latents.to(torch.float32)
    dt = sigma_next - sigma
    prev = sample + dt * noise_pred
    return prev.to(noise_pred.dtype)


def diffuse(w: dict, dims: DiTDims, latents, prompt_embeds, neg_prompt_embeds, sigmas: np.ndarray, img_shape,
            true_cfg_scale: float = 4.0, image_latents=None):
    """QwenImagePipeline.diffuse (pipeline_qwen_image.py:530-586).  `sigmas` has N+1 entries
    (last = 0); timesteps = sigmas[:-1]*1000 in fp32.  With `image_latents` [B,S2,64] it is
    QwenImageEditPipeline.diffuse (pipeline_qwen_image_edit.py:574-639): the condition latents are appended on the
    sequence axis every step (:600-602), `img_shape` lists both grids and only the first S1 rows of the prediction are
    kept (:617,632)."""
    sig = torch.from_numpy(np.asarray(sigmas, dtype=np.float32))
    timesteps = sig[:-1] * 1000.0
    do_cfg = neg_prompt_embeds is not None
    for i, t in enumerate(timesteps):
        timestep = t.expand(latents.shape[0]).to(dtype=latents.dtype)  # (:552) rounds t to bf16 in bf16 mode
        x_in = latents if image_latents is None else torch.cat([latents, image_latents], dim=1)
        noise = model_forward(w, dims, x_in, prompt_embeds, timestep / 1000, img_shape)[:, : latents.shape[1]]
        if do_cfg:
            neg = model_forward(w, dims, x_in, neg_prompt_embeds, timestep / 1000, img_shape)[:, : latents.shape[1]]
            noise = cfg_combine(noise, neg, true_cfg_scale)
        latents = euler_step(noise, latents, sig[i], sig[i + 1])
    return latents


def pack_latents(latents, batch_size, num_channels_latents, height, width):
    """QwenImagePipeline._pack_latents (pipeline_qwen_image.py:435-441)."""
    latents = latents.view(batch_size, num_channels_latents, height // 2, 2, width // 2, 2)
    latents = latents.permute(0, 2, 4, 1, 3, 5)
    return latents.reshape(batch_size, (height // 2) * (width // 2), num_channels_latents * 4)


# --------------------------------------------------------------------------------------
# helpers for tests
# --------------------------------------------------------------------------------------
def rel_fro(a: torch.Tensor, b: torch.Tensor) -> float:
    """relative Frobenius error ||a-b|| / ||b|| in fp64 (SURVEY.md §8d parity metric)."""
    a = a.double()
    b = b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def cast_weights(w: dict, dtype) -> dict:
    return {k: v.to(dtype) for k, v in w.items()}


def flops_per_forward(dims: DiTDims, s_img: int, t: int) -> float:
    """Algorithmic FLOPs per image per forward."""
    D = dims.dim
    S = s_img + t
    return dims.num_layers * (24 * S * D * D + 4 * S * S * D) + 2 * s_img * dims.in_channels * D * 2 + 2 * t * dims.joint_dim * D
