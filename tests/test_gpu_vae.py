"""GPU (-m gpu): the native VAE decode (csrc/qimg_vae.cu through the C-ABI) against the fp32 oracle
(oracle/vae_oracle.py, pinned to the unmodified reference by oracle/make_golden_vae.py) and the reference's golden images.

Tolerances.  The reference computes these convolutions on the GPU with TF32 inputs / fp32 accumulation (cuDNN,
`torch.backends.cudnn.allow_tf32` default) — the arithmetic the native tcgen05 kind::tf32 kernels use — and the goldens are
the reference in full fp32 on the CPU.  A CPU emulation of TF32 input rounding through the whole decoder sits at 1.2e-3
relative Frobenius / 3e-3 max abs of the fp32 image; the bars below leave 3x headroom over that:
  one convolution / GEMM    <= 2e-3 relative Frobenius of the fp32 result  (TF32: 10 mantissa bits)
  whole decode              <= 4e-3 relative Frobenius, <= 1.5e-2 max abs, and after the pipeline's uint8 quantisation
                            >= 85 % of the pixels identical, none off by more than one level
  row kernels (norm, upsample, softmax, transpose, post-quant, conv_out)   fp32 round-off (<= 2e-5)
"""
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import vae_oracle
from vllm_omni_b200 import lib as q
from vllm_omni_b200 import synthetic
from vllm_omni_b200.diffusion.models.qwen_image.vae_decoder import B200AutoencoderKLQwenImage, B200VaeDecoder, _pack3x3

pytestmark = pytest.mark.gpu
dev = "cuda"
TOL_TF32 = 2e-3


def gen(seed):
    return torch.Generator().manual_seed(seed)


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def nhwc(t):  # [N, C, H, W] -> [N, H, W, C] contiguous
    return t.permute(0, 2, 3, 1).contiguous()


@pytest.fixture(params=[0, 1], ids=["shared-dy-taps", "box-per-tap"])
def conv_variant(request):
    """Both implicit-GEMM kernels: the default (one TMA box per horizontal tap shared by the three vertical taps, 16 x 16
    pixel patches) and the first version (one box per tap, 16 x 8 patches)."""
    q.set_vae_conv_variant(request.param)
    yield request.param
    q.set_vae_conv_variant(0)


@pytest.mark.parametrize("N,H,W,cin,cout", [(1, 8, 16, 32, 128), (2, 19, 23, 96, 96), (1, 40, 36, 384, 192), (1, 16, 32, 192, 384),
                                            (4, 128, 160, 32, 96), (4, 128, 160, 32, 192), (4, 122, 166, 32, 128)],
                         ids=["one-tile", "ragged-96", "384to192", "192to384", "256px-tiles-n96", "256px-tiles-n192-one-acc-stage",
                              "256px-tiles-n128-ragged"])
def test_conv3x3_tf32_vs_fp32(N, H, W, cin, cout, conv_variant):
    g = gen(H * 1000 + W)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    res = torch.randn(N, cout, H, W, generator=g)
    want = F.conv2d(x, w, b, padding=1)
    got = q.conv2d_nhwc_tf32(nhwc(x).to(dev), _pack3x3(w).to(dev), b.to(dev), 9, cout)
    assert rel(got.cpu().permute(0, 3, 1, 2), want) < TOL_TF32
    got = q.conv2d_nhwc_tf32(nhwc(x).to(dev), _pack3x3(w).to(dev), b.to(dev), 9, cout, res=nhwc(res).to(dev))
    assert rel(got.cpu().permute(0, 3, 1, 2), want + res) < TOL_TF32


def test_conv3x3_padding_is_zero_fill_at_every_border(conv_variant):
    """An all-ones image through an all-ones 3x3 kernel counts the taps that meet data: 4 in the corners, 6 on the edges,
    9 inside — exact in TF32, so any tap read from outside the image (or from the neighbouring image of the batch) shows."""
    for N, H, W, C in ((2, 11, 21, 32), (4, 120, 170, 32)):  # the second shape is large enough for 16 x 16 patches per CTA
        x = torch.ones(N, H, W, C, device=dev)
        w = torch.zeros(4, 9 * C, device=dev)
        w[:, ::C] = 1.0  # channel 0 of every tap
        got = q.conv2d_nhwc_tf32(x, w, None, 9, 4).cpu()
        want = F.conv2d(torch.ones(N, 1, H, W), torch.ones(1, 1, 3, 3), padding=1).permute(0, 2, 3, 1).expand(N, H, W, 4)
        assert torch.equal(got, want)
    # a kernel that distinguishes the taps: tap t weighs 2^t, so the sum names exactly which taps met data
    x = torch.ones(1, 24, 40, 32, device=dev)
    w = torch.zeros(4, 9 * 32, device=dev)
    for t in range(9):
        w[:, t * 32] = 2.0 ** t
    got = q.conv2d_nhwc_tf32(x, w, None, 9, 4).cpu()
    kern = (2.0 ** torch.arange(9.0)).view(1, 1, 3, 3)
    want = F.conv2d(torch.ones(1, 1, 24, 40), kern, padding=1).permute(0, 2, 3, 1).expand(1, 24, 40, 4)
    assert torch.equal(got, want)


@pytest.mark.parametrize("P_h,P_w,K,Nout", [(8, 16, 384, 1152), (18, 22, 396, 384), (16, 16, 64, 256), (64, 64, 200, 4096)],
                         ids=["qkv", "k-tail-396", "small", "scores-256px-tiles"])
def test_gemm_mode_1x1_vs_fp32(P_h, P_w, K, Nout, conv_variant):
    """taps = 1: out[pixels, Nout] = x[pixels, K] w[Nout, K]^T, including a K that is no multiple of the 32-float K block
    (the P*V product of a 18 x 22 latent: 396 keys) and strided operands (channel slices of a wider buffer)."""
    g = gen(K)
    P = P_h * P_w
    xbuf = torch.randn(P, K + 8, generator=g)
    wbuf = torch.randn(Nout, K + 4, generator=g) / K ** 0.5
    res = torch.randn(P, Nout, generator=g)
    want = xbuf[:, :K] @ wbuf[:, :K].T + res
    xd = xbuf.to(dev).view(1, P_h, P_w, K + 8)[..., :K]
    got = q.conv2d_nhwc_tf32(xd, wbuf.to(dev)[:, :K], None, 1, Nout, res=res.to(dev).view(1, P_h, P_w, Nout), cin=K)
    assert rel(got.cpu().view(P, Nout), want) < TOL_TF32


def test_row_kernels_match_torch():
    g = gen(5)
    for C in (96, 192, 384):
        x = torch.randn(3, 9, 17, C, generator=g) * 3
        gamma = 1 + 0.1 * torch.randn(C, generator=g)
        for silu in (False, True):
            want = F.normalize(x, dim=-1) * C ** 0.5 * gamma
            want = F.silu(want) if silu else want
            got = q.vae_rms_act(x.to(dev), gamma.to(dev), silu).cpu()
            assert (got - want).abs().max().item() < 2e-5
    x = torch.randn(2, 5, 7, 96, generator=g)
    want = F.interpolate(x.permute(0, 3, 1, 2), scale_factor=(2.0, 2.0), mode="nearest-exact").permute(0, 2, 3, 1)
    assert torch.equal(q.vae_upsample2x(x.to(dev)).cpu(), want)
    for cols in (396, 397, 30000):  # float4 rows in shared memory, scalar rows, rows too long for shared memory
        s = torch.randn(37, cols, generator=g) * 20
        want = torch.softmax(s * 0.051, dim=-1)
        assert (q.vae_softmax_rows(s.to(dev), 0.051).cpu() - want).abs().max().item() < 2e-6
    t = torch.randn(70, 1152, generator=g)
    assert torch.equal(q.vae_transpose(t.to(dev)[:, 768:]).cpu(), t[:, 768:].T.contiguous())
    z = torch.randn(2, 16, 9, 13, generator=g)
    w, b = torch.randn(16, 16, generator=g) / 4, torch.randn(16, generator=g)
    got = q.vae_post_quant(z.to(dev), w.to(dev), b.to(dev)).cpu()
    want = F.conv2d(z, w.view(16, 16, 1, 1), b).permute(0, 2, 3, 1)
    assert (got[..., :16] - want).abs().max().item() < 2e-5 and not got[..., 16:].any()
    y = torch.randn(2, 10, 37, 96, generator=g)
    w3, b3 = torch.randn(3, 96, 3, 3, generator=g) / 30, torch.randn(3, generator=g) * 0.1
    want = F.conv2d(y.permute(0, 3, 1, 2), w3, b3, padding=1).clamp(-1, 1)
    got = q.vae_conv_out(y.to(dev), w3.permute(0, 2, 3, 1).contiguous().to(dev), b3.to(dev)).cpu()
    assert (got - want).abs().max().item() < 2e-5


def _check_image(got, want):
    d = (got - want).abs()
    u8 = lambda t: ((t * 0.5 + 0.5) * 255).round().clamp(0, 255)  # noqa: E731  (the post-process of pipeline_qwen_image.py:40-60)
    lv = (u8(got) - u8(want)).abs()
    stats = dict(rel=rel(got, want), max_abs=d.max().item(), same=(lv == 0).float().mean().item(), max_level=lv.max().item())
    assert stats["rel"] < 4e-3 and stats["max_abs"] < 1.5e-2 and stats["same"] > 0.85 and stats["max_level"] <= 1, stats
    return stats


@pytest.mark.parametrize("name", ["vae_decode_ragged", "vae_decode_256px"])
def test_decode_matches_reference_golden(golden_dir, name):
    gold = torch.load(os.path.join(golden_dir, name + ".pt"))
    W = synthetic.synthetic_vae_decoder_weights(seed=gold["wseed"])
    assert abs(sum(v.double().abs().sum() for v in W.values()).item() - gold["weights_checksum"]) < 1e-6 * gold["weights_checksum"]
    vae = B200VaeDecoder(W, device=dev)
    n0 = q.launch_count()
    img = vae.decode(gold["z"].to(dev), return_dict=False)[0]
    assert q.launch_count() - n0 > 60  # the sm_100a kernels ran (no torch fallback exists)
    assert img.shape == gold["image"].shape and img.dtype == torch.float32
    _check_image(img.cpu(), gold["image"])


def test_decode_batch_items_are_independent_and_banded_attention_is_identical(monkeypatch):
    W = synthetic.synthetic_vae_decoder_weights(seed=5)
    vae = B200VaeDecoder(W, device=dev)
    z = torch.randn(3, 16, 1, 16, 24, generator=gen(21)).to(dev)
    full = vae.decode(z, return_dict=False)[0]
    for i in range(3):
        assert torch.equal(vae.decode(z[i:i + 1], return_dict=False)[0], full[i:i + 1])
    import vllm_omni_b200.diffusion.models.qwen_image.vae_decoder as V
    monkeypatch.setattr(V, "SCORES_BYTES_MAX", 4 * (16 * 24) * 24 * 8)  # 8 image rows per score band, two bands
    assert torch.equal(vae.decode(z, return_dict=False)[0], full)


def test_decode_to_uint8_is_the_reference_post_process_of_the_decoded_image():
    """decode_to_uint8 == VaeImageProcessor.postprocess arithmetic (pipeline_qwen_image.py:40-60) applied to decode():
    bit-exact, and the registry's post-process function turns both into the same PIL images."""
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import get_qwen_image_post_process_func
    vae = B200VaeDecoder(synthetic.synthetic_vae_decoder_weights(seed=8), device=dev)
    z = torch.randn(2, 16, 1, 9 * 2, 16, generator=gen(51)).to(dev)
    img = vae.decode(z, return_dict=False)[0][:, :, 0]
    u8 = vae.decode_to_uint8(z)
    want = ((img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).round().to(torch.uint8)
    assert u8.shape == (2, 144, 128, 3) and torch.equal(u8, want)
    post = get_qwen_image_post_process_func(None)
    a, b = post(img), post(u8)
    assert len(a) == 2 and a[0].size == (128, 144) and all(x.tobytes() == y.tobytes() for x, y in zip(a, b))


def test_decode_1024px_against_oracle():
    """BASELINE configs[1] geometry (1024 x 1024: 128 x 128 latent grid, 16384 attention positions) against the fp32 oracle."""
    W = synthetic.synthetic_vae_decoder_weights(seed=6)
    z = torch.randn(1, 16, 1, 128, 128, generator=gen(31))
    img = B200VaeDecoder(W, device=dev).decode(z.to(dev), return_dict=False)[0]
    assert img.shape == (1, 3, 1, 1024, 1024)
    want = vae_oracle.vae_decode(z, W)
    _check_image(img.cpu(), want)


def test_pipeline_decodes_through_the_native_vae():
    """QwenImagePipeline.forward's post-step (pipeline_qwen_image.py:736-747) with the native decoder injected as `vae`:
    unpack + de-normalise + decode, against the oracle on the same latents."""
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    W = synthetic.synthetic_vae_decoder_weights(seed=7)
    vae = B200VaeDecoder(W, device=dev)
    h, w = 16, 32  # latent grid -> 128 x 256 px
    lat = torch.randn(2, (h // 2) * (w // 2), 64, generator=gen(41))
    unpacked = QwenImagePipeline._unpack_latents(lat, 8 * h, 8 * w, 8)
    mean = torch.tensor(vae.config.latents_mean).view(1, 16, 1, 1, 1)
    std = 1.0 / torch.tensor(vae.config.latents_std).view(1, 16, 1, 1, 1)
    want = vae_oracle.vae_decode(unpacked / std + mean, W)[:, :, 0]
    got = QwenImagePipeline.decode_latents(vae, lat.to(dev), 8 * h, 8 * w)
    _check_image(got.cpu(), want)


def test_forward_produces_an_image_end_to_end():
    """`QwenImagePipeline.forward` from prompt embeddings to pixels with every stage native: DiT denoise (engine) -> unpack
    -> de-normalise -> VAE decode -> (optionally) the uint8 post-process.  The image equals the fp32 oracle's decode of the
    latents the same request returns with output_type="latent" (the DiT side has its own goldens)."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline, get_qwen_image_post_process_func
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}))
    W = synthetic.synthetic_vae_decoder_weights(seed=9)
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint),
                                     vae=B200VaeDecoder(W, device=dev))
    finally:
        torch.set_default_dtype(torch.float32)
    pipe.transformer.load_weights(dict(synthetic.synthetic_weights(L, seed=43, norm_jitter=0.1, num_heads=H, joint_dim=joint)).items())
    g = gen(44)
    B, hh, ww, T = 2, 8, 16, 24   # 128 x 256 px: latent grid 16 x 32
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    pe = torch.randn(B, T, joint, generator=g).bfloat16()
    kw = dict(prompt_embeds=pe, latents=lat, height=hh * 16, width=ww * 16, num_inference_steps=3, true_cfg_scale=1.0)
    final = pipe.forward(OmniDiffusionRequest(output_type="latent", **kw)).output
    out = pipe.forward(OmniDiffusionRequest(output_type="pil", **kw))
    assert out.error is None and out.output.shape == (B, 3, hh * 16, ww * 16) and out.output.dtype == torch.float32
    unpacked = QwenImagePipeline._unpack_latents(final.cpu(), hh * 16, ww * 16, 8).float()
    mean = torch.tensor(pipe.vae.config.latents_mean).view(1, 16, 1, 1, 1)
    std = 1.0 / torch.tensor(pipe.vae.config.latents_std).view(1, 16, 1, 1, 1)
    want = vae_oracle.vae_decode(unpacked / std + mean, W)[:, :, 0]
    _check_image(out.output.cpu(), want)
    pipe.uint8_output = True
    u8 = pipe.forward(OmniDiffusionRequest(output_type="pil", **kw)).output
    assert u8.dtype == torch.uint8 and u8.shape == (B, hh * 16, ww * 16, 3)
    post = get_qwen_image_post_process_func(od)
    imgs, imgs_u8 = post(out.output), post(u8)
    assert len(imgs) == B and imgs[0].size == (ww * 16, hh * 16) and all(a.tobytes() == b.tobytes() for a, b in zip(imgs, imgs_u8))


# ---- encode side (the edit pipelines' condition image) -----------------------------------------------------------------
@pytest.mark.parametrize("N,H,W,cin,cout", [(1, 16, 32, 32, 96), (2, 36, 44, 96, 96), (1, 72, 88, 192, 192)],
                         ids=["one-tile", "ragged", "192"])
def test_conv3x3_stride2_vs_fp32(N, H, W, cin, cout):
    """The encoder's resamplers: ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2) (autoencoder_kl_qwenimage.py:157-161); the
    stride-2 gather is the TMA descriptor's element stride."""
    g = gen(H + W)
    x = torch.randn(N, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (9 * cin) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    want = F.conv2d(F.pad(x, (0, 1, 0, 1)), w, b, stride=2)
    got = q.conv2d_down2_nhwc_tf32(nhwc(x).to(dev), _pack3x3(w).to(dev), b.to(dev), cout)
    assert got.shape == (N, H // 2, W // 2, cout)
    assert rel(got.cpu().permute(0, 3, 1, 2), want) < TOL_TF32
    # which input pixels a tap meets: weights 2^tap on channel 0, input = a per-pixel code that is exact in TF32
    xs = torch.zeros(1, 32, 16, 32)
    xs[0, 0] = (torch.arange(16).view(16, 1) * 32 + torch.arange(32).view(1, 32)).float()  # <= 511: exact in TF32
    wt = torch.zeros(4, 32, 3, 3)
    for t in range(9):
        wt[:, 0, t // 3, t % 3] = 2.0 ** t
    want = F.conv2d(F.pad(xs, (0, 1, 0, 1)), wt, None, stride=2).permute(0, 2, 3, 1)
    assert torch.equal(q.conv2d_down2_nhwc_tf32(nhwc(xs).to(dev), _pack3x3(wt).to(dev), None, 4).cpu(), want)


def test_image_to_nhwc():
    img = torch.randn(2, 3, 9, 13, generator=gen(3))
    got = q.vae_image_to_nhwc(img.to(dev)).cpu()
    assert torch.equal(got[..., :3], img.permute(0, 2, 3, 1)) and not got[..., 3:].any()


def test_encode_matches_reference_golden(golden_dir):
    """Posterior parameters of the native encoder against the unmodified reference VAE's `encode` (fp32, CPU): TF32 level."""
    gold = torch.load(os.path.join(golden_dir, "vae_encode_ragged.pt"))
    W = {**synthetic.synthetic_vae_decoder_weights(seed=1), **synthetic.synthetic_vae_encoder_weights(seed=gold["wseed"])}
    vae = B200AutoencoderKLQwenImage(W, device=dev)
    n0 = q.launch_count()
    dist = vae.encode(gold["x"].to(dev)).latent_dist
    assert q.launch_count() - n0 > 50
    assert dist.parameters.shape == gold["params"].shape
    assert rel(dist.parameters.cpu(), gold["params"]) < 4e-3
    assert rel(dist.mode().cpu(), gold["params"][:, :16]) < 4e-3
    # batch items are independent
    assert torch.equal(vae.encode(gold["x"][1:].to(dev)).latent_dist.parameters, dist.parameters[1:])


def test_edit_request_with_a_pixel_image_equals_the_same_request_with_its_latents():
    """QwenImageEditPipeline.forward with req.extra['image'] (native VAE encode -> normalise -> pack) is bit-identical to the
    request that carries the condition latents the encoder produces; the latents match the oracle's encode of the image."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image_edit import QwenImageEditPipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}))
    W = {**synthetic.synthetic_vae_decoder_weights(seed=1), **synthetic.synthetic_vae_encoder_weights(seed=2)}
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(dev):
            pipe = QwenImageEditPipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint),
                                         vae=B200AutoencoderKLQwenImage(W, device=dev))
    finally:
        torch.set_default_dtype(torch.float32)
    pipe.transformer.load_weights(dict(synthetic.synthetic_weights(L, seed=45, norm_jitter=0.1, num_heads=H, joint_dim=joint)).items())
    g = gen(46)
    B, hh, ww, T = 1, 8, 16, 24
    img = torch.rand(1, 3, 64, 128, generator=g) * 2 - 1      # condition image: latent 8 x 16 -> 4 x 8 patches
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    pe = torch.randn(B, T, joint, generator=g).bfloat16()
    kw = dict(prompt_embeds=pe, latents=lat, height=hh * 16, width=ww * 16, num_inference_steps=2, true_cfg_scale=1.0,
              output_type="latent")
    out_px = pipe.forward(OmniDiffusionRequest(extra={"image": img}, **kw))
    assert out_px.error is None
    il, grid = pipe._encode_vae_image(img)
    assert grid == (4, 8) and il.shape == (1, 32, 64)
    out_lat = pipe.forward(OmniDiffusionRequest(extra={"image_latents": il, "image_latent_grid": grid}, **kw))
    assert torch.equal(out_px.output, out_lat.output)
    params = vae_oracle.vae_encode(img.unsqueeze(2), W)
    mean = torch.tensor(pipe.vae.config.latents_mean).view(1, 16, 1, 1, 1)
    std = torch.tensor(pipe.vae.config.latents_std).view(1, 16, 1, 1, 1)
    want = QwenImageEditPipeline._pack_latents((params[:, :16] - mean) / std, 1, 16, 8, 16)
    assert rel(il.float().cpu(), want) < 6e-3  # TF32 convolutions + the bf16 rounding of the packed latents
