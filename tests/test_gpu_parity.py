"""GPU (-m gpu): parity of the sm_100a path against the oracle and the reference golden vectors.
Everything calls through the C-ABI library (vllm_omni_b200.lib / the engine); tolerances follow SURVEY §8d:
  (i)  fused kernel vs fp32 restatement        <= 2^-7 relative Frobenius
  (ii) one block / L=2 model vs reference-bf16 <= 1e-2
  (iii) err(native, fp32) <= err(reference-bf16, fp32) + 1e-2
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import qwen_image_oracle as O
from vllm_omni_b200 import lib as q
from vllm_omni_b200 import synthetic

pytestmark = pytest.mark.gpu
dev = "cuda"
bf = torch.bfloat16
TOL_KERNEL = 2.0 ** -7


def gen(seed):
    return torch.Generator().manual_seed(seed)


@pytest.fixture(params=[0, 1], ids=["cta1", "cta2pair"])
def gemm_mode(request):
    """Run a GEMM-bearing test under both tcgen05 tile modes (cta_group::1 and the cta_group::2 CTA pair)."""
    prev = q.get_gemm_mode()
    q.set_gemm_mode(request.param)
    yield request.param
    q.set_gemm_mode(prev)


def make_model(L, H, joint, seed, norm_jitter=0.1):
    from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            m = QwenImageTransformer2DModel(num_layers=L, num_attention_heads=H, joint_attention_dim=joint)
    finally:
        torch.set_default_dtype(torch.float32)
    m.load_weights(synthetic.split_qkv_checkpoint_names(
        synthetic.synthetic_weights(L, seed=seed, norm_jitter=norm_jitter, num_heads=H, joint_dim=joint)))
    return m


def test_library_loaded_and_device():
    assert q.device_check() >= 100  # B200: 148 SMs
    assert os.path.exists(q.LIB_PATH)


@pytest.fixture(params=[(4, 0), (6, 0), (12, 0), (14, 0), (4, 1), (6, 1)],
                ids=["exact", "fast", "exact-poly25", "fast-poly25", "exact-1tile", "fast-1tile"])
def fmha_mode(request):
    """Both shipped attention pipelines (exact = per-tile maximum first; fast = running reference maximum + overflow
    guard), with and without the 25 % FMA-pipe polynomial share, and both query tilings (a CTA per pair of query tiles /
    per single tile — forced here, chosen by grid size in production)."""
    prev = q.get_fmha_mode()
    q.set_fmha_mode(request.param[0])
    q.set_fmha_single_tile(request.param[1])
    yield request.param[0]
    q.set_fmha_mode(prev)
    q.set_fmha_single_tile(-1)


def test_umma_probe_all_operand_paths():
    g = gen(0)
    A = torch.randn(128, 128, generator=g).bfloat16()
    Bm = torch.randn(128, 128, generator=g).bfloat16()
    for mode in (0, 1, 2):
        D = q.umma_probe(A.to(dev), Bm.to(dev), mode).cpu()
        ref = A.float() @ (Bm.float().T if mode == 0 else Bm.float())
        assert O.rel_fro(D, ref) < 1e-5, mode


@pytest.mark.parametrize("B,S,D", [(2, 40, 256), (1, 33, 3072), (3, 7, 512), (1, 1, 128)])
def test_ln_modulate(B, S, D):
    g = gen(1)
    x = (torch.randn(B, S, D, generator=g) * 2 + 0.3).bfloat16()
    mod = (torch.randn(B, 3 * D, generator=g) * 0.5).bfloat16()
    ref, _ = O.ada_layer_norm(x, mod, 1e-6)
    md = mod.to(dev)
    y = q.ln_modulate(x.view(-1, D).to(dev), md[:, :D], md[:, D:2 * D], S, 3 * D).view(B, S, D).cpu()
    assert O.rel_fro(y, ref) < 1e-3
    ref32, _ = O.ada_layer_norm(x.float(), mod.float(), 1e-6)
    assert O.rel_fro(y, ref32) < TOL_KERNEL


def test_adalayernorm_token_index_matches_reference_golden(golden_dir):
    """Per-token modulation select (`zero_cond_t` models; AdaLayerNorm with `index`, layers/adalayernorm.py:31-54): the fused
    kernel + the gate gather against the unmodified reference layer's outputs, through the CustomOp plug-in."""
    from vllm_omni_b200.diffusion.layers.adalayernorm import AdaLayerNorm
    gold = torch.load(os.path.join(golden_dir, "adaln_index.pt"))
    for name, c in gold.items():
        D = c["x"].shape[-1]
        y, gate = AdaLayerNorm(D, eps=1e-6).forward_cuda(c["x"].to(dev), c["mod"].to(dev), c["index"].to(dev))
        assert torch.equal(gate.cpu(), c["gate"]), name                       # a gather: exact
        assert O.rel_fro(y.cpu(), c["y"]) < 1e-3, name                        # vs the reference's bf16 output
        assert O.rel_fro(y.cpu(), c["y_fp32"]) < TOL_KERNEL, name             # vs fp32
        # all-zero index == the un-indexed path on the first half of the modulation rows
        B = c["x"].shape[0]
        y0, g0 = AdaLayerNorm(D, eps=1e-6).forward_cuda(c["x"].to(dev), c["mod"].to(dev), torch.zeros_like(c["index"]).to(dev))
        y1, g1 = AdaLayerNorm(D, eps=1e-6).forward_cuda(c["x"].to(dev), c["mod"][:B].to(dev))
        assert torch.equal(y0, y1) and torch.equal(g0, g1.expand_as(g0))


def test_adalayernorm_custom_op_plugin():
    from vllm_omni_b200.diffusion.layers.adalayernorm import AdaLayerNorm
    g = gen(2)
    x = torch.randn(2, 19, 256, generator=g).bfloat16()
    mod = torch.randn(2, 768, generator=g).bfloat16()
    y, gate = AdaLayerNorm(256)(x.to(dev), mod.to(dev))
    ref, rg = O.ada_layer_norm(x, mod, 1e-6)
    assert O.rel_fro(y.cpu(), ref) < 1e-3 and torch.equal(gate.cpu(), rg)


def test_rms_norm_and_gate_residual():
    g = gen(3)
    x = torch.randn(37, 3584, generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(3584, generator=g)).bfloat16()
    assert O.rel_fro(q.rms_norm(x.to(dev), w.to(dev)).cpu(), O.rms_norm(x, w, 1e-6)) < 1e-3
    x = torch.randn(2, 19, 256, generator=g).bfloat16()
    y = torch.randn(2, 19, 256, generator=g).bfloat16()
    gate = torch.randn(2, 256, generator=g).bfloat16()
    xd = x.view(-1, 256).to(dev).clone()
    q.gate_residual(xd, y.view(-1, 256).to(dev), gate.to(dev), 19, 256)
    assert torch.equal(xd.cpu().view(2, 19, 256), x + gate[:, None, :] * y)


def test_gate_residual_bias_tp_form():
    g = gen(31)
    x = torch.randn(2, 19, 256, generator=g).bfloat16()
    y = torch.randn(2, 19, 256, generator=g).bfloat16()
    bias = torch.randn(256, generator=g).bfloat16()
    gate = torch.randn(2, 256, generator=g).bfloat16()
    xd, yd, bd, gd = x.view(-1, 256).to(dev).clone(), y.view(-1, 256).to(dev), bias.to(dev), gate.to(dev)  # keep alive
    q.check(q.load().qimg_gate_residual_bias(xd.data_ptr(), yd.data_ptr(), bd.data_ptr(), gd.data_ptr(), 38, 256, 19, 256,
                                            q.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(xd.cpu().view(2, 19, 256), x + gate[:, None, :] * (y + bias))


def test_row_parallel_partial_sums_match_full_gemm(gemm_mode):
    """The TP engine's row-parallel linears: sum over K-shards of (A_shard @ W_shard^T) with a zero bias equals
    the full linear (bf16 partial sums, so only to bf16 accuracy) — what the all-reduce reconstructs."""
    g = gen(32)
    M, N, K, P = 300, 512, 1024, 2
    A = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) / 32).bfloat16()
    zero = torch.zeros(N, dtype=bf, device=dev)
    parts = [q.linear(A[:, r * K // P:(r + 1) * K // P].contiguous().to(dev), W[:, r * K // P:(r + 1) * K // P].contiguous().to(dev), zero)
             for r in range(P)]
    total = (parts[0].float() + parts[1].float()).cpu()
    assert O.rel_fro(total, A.float() @ W.float().T) < TOL_KERNEL


@pytest.mark.parametrize("M,N,K,act", [(1, 1536, 256, True), (4, 6216, 3072, True), (3, 512, 256, False), (11, 640, 512, True)])
def test_linear_small_m(M, N, K, act):
    g = gen(4)
    x = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
    b = torch.randn(N, generator=g).bfloat16()
    ref = F.linear(F.silu(x) if act else x, W, b)
    assert O.rel_fro(q.linear_small_m(x.to(dev), W.to(dev), b.to(dev), act).cpu(), ref) < 1e-3


def test_timestep_sinusoid():
    t = torch.tensor([1000.0, 731.0, 20.5, 0.0]).bfloat16() / 1000
    assert O.rel_fro(q.timestep_sinusoid(t.to(dev)).cpu(), O.timestep_sinusoid(t.float()).bfloat16()) < 1e-3


@pytest.mark.parametrize("cfg", [True, False])
def test_cfg_euler_step_matches_oracle(cfg):
    g = gen(5)
    pos, neg, lat = (torch.randn(2, 50, 64, generator=g).bfloat16() for _ in range(3))
    sig, sig_n = torch.tensor(0.9137), torch.tensor(0.8811)
    noise = O.cfg_combine(pos, neg, 4.0) if cfg else pos
    ref = O.euler_step(noise, lat, sig, sig_n)
    ld = lat.to(dev).clone()
    q.cfg_euler_step(pos.to(dev), neg.to(dev) if cfg else None, ld, 4.0, float(sig), float(sig_n))
    got = ld.cpu()
    # elementwise chain is rounded at the reference's points: allow 1 bf16 ulp on a few elements (norm reduction order)
    assert O.rel_fro(got, ref) < 2e-3
    assert (got.float() - ref.float()).abs().max() <= 2.0 ** -6 * ref.float().abs().max()


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 768, 256), (1024, 3072, 3072), (77, 64, 512), (512, 12288, 3072), (5, 384, 128)])
def test_gemm_bias_and_gelu(M, N, K, gemm_mode):
    g = gen(6)
    x = torch.randn(M, K, generator=g).bfloat16()
    W = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16()
    b = torch.randn(N, generator=g).bfloat16()
    ref = F.linear(x.float(), W.float(), b.float())
    assert O.rel_fro(q.linear(x.to(dev), W.to(dev), b.to(dev)).cpu(), ref) < TOL_KERNEL
    if N > 64:
        y = q.linear(x.to(dev), W.to(dev), b.to(dev), q.EPI_BIAS_GELU).cpu()
        assert O.rel_fro(y, F.gelu(ref, approximate="tanh")) < TOL_KERNEL
        assert O.rel_fro(y, F.gelu(F.linear(x, W, b), approximate="tanh")) < 5e-3  # reference bf16 op order


def test_gemm_tile_modes_bit_identical():
    """CTA-pair (256-row) and single-CTA (128-row) tiles run the same K loop per output element: bit-identical outputs.
    The launcher relies on it when it picks the tile shape by grid size (mode 2), and so does the bit-identity of the
    sequence-parallel engine (own-row GEMMs with small M may use the other tile shape than the single-GPU GEMM)."""
    g = gen(60)
    prev = q.get_gemm_mode()
    try:
        for (M, N, K) in ((1088, 3072, 3072), (300, 768, 1024), (2176, 12288, 3072)):
            x = torch.randn(M, K, generator=g).bfloat16().to(dev)
            W = (torch.randn(N, K, generator=g) * K ** -0.5).bfloat16().to(dev)
            b = torch.randn(N, generator=g).bfloat16().to(dev)
            outs = []
            for mode in (0, 1, 2):
                q.set_gemm_mode(mode)
                outs.append(q.linear(x, W, b, q.EPI_BIAS_GELU).clone())
            assert torch.equal(outs[0], outs[1]) and torch.equal(outs[1], outs[2]), (M, N, K)
    finally:
        q.set_gemm_mode(prev)


def test_gemm_grouped_gate_residual(gemm_mode):
    g = gen(7)
    Mi, Mt, D, K = 2 * 200, 2 * 24, 256, 1024
    xi, xt = torch.randn(Mi, D, generator=g).bfloat16(), torch.randn(Mt, D, generator=g).bfloat16()
    ai, at = torch.randn(Mi, K, generator=g).bfloat16(), torch.randn(Mt, K, generator=g).bfloat16()
    Wi, Wt = (torch.randn(D, K, generator=g) / 32).bfloat16(), (torch.randn(D, K, generator=g) / 32).bfloat16()
    bi, bt = torch.randn(D, generator=g).bfloat16(), torch.randn(D, generator=g).bfloat16()
    gi, gt = torch.randn(2, D, generator=g).bfloat16(), torch.randn(2, D, generator=g).bfloat16()
    ref_i = xi.view(2, 200, D) + gi[:, None] * F.linear(ai, Wi, bi).view(2, 200, D)
    ref_t = xt.view(2, 24, D) + gt[:, None] * F.linear(at, Wt, bt).view(2, 24, D)
    xid, xtd = xi.to(dev).clone(), xt.to(dev).clone()
    k = [t.to(dev) for t in (ai, Wi, bi, gi, at, Wt, bt, gt)]
    p0 = q.GemmProblem(A=k[0].data_ptr(), W=k[1].data_ptr(), bias=k[2].data_ptr(), M=Mi, N=D, K=K, rows_per_batch=200,
                       out=xid.data_ptr(), ldo=D, gate=k[3].data_ptr(), gate_stride=D)
    p1 = q.GemmProblem(A=k[4].data_ptr(), W=k[5].data_ptr(), bias=k[6].data_ptr(), M=Mt, N=D, K=K, rows_per_batch=24,
                       out=xtd.data_ptr(), ldo=D, gate=k[7].data_ptr(), gate_stride=D)
    q.gemm([p0, p1], q.EPI_BIAS_GATE_RES)
    assert O.rel_fro(xid.cpu().view(2, 200, D), ref_i) < 2e-3
    assert O.rel_fro(xtd.cpu().view(2, 24, D), ref_t) < 2e-3


def _qkv_reference(w, p, img, txt, H, rope):
    def proj(x, wn, bn):
        return (t.unflatten(-1, (H, -1)) for t in F.linear(x, w[p + wn], w[p + bn]).chunk(3, dim=-1))
    iq, ik, iv = proj(img, "attn.to_qkv.weight", "attn.to_qkv.bias")
    tq, tk, tv = proj(txt, "attn.add_kv_proj.weight", "attn.add_kv_proj.bias")
    iq, ik = O.rms_norm(iq, w[p + "attn.norm_q.weight"], 1e-6), O.rms_norm(ik, w[p + "attn.norm_k.weight"], 1e-6)
    tq, tk = O.rms_norm(tq, w[p + "attn.norm_added_q.weight"], 1e-6), O.rms_norm(tk, w[p + "attn.norm_added_k.weight"], 1e-6)
    r = [t.to(bf) for t in rope]
    iq, ik = O.apply_rope_interleaved(iq, r[0], r[1]), O.apply_rope_interleaved(ik, r[0], r[1])
    tq, tk = O.apply_rope_interleaved(tq, r[2], r[3]), O.apply_rope_interleaved(tk, r[2], r[3])
    return torch.cat([tq, iq], 1), torch.cat([tk, ik], 1), torch.cat([tv, iv], 1)


@pytest.mark.parametrize("B,h,wd,T,H", [(1, 8, 16, 128, 2), (2, 10, 9, 37, 2), (1, 32, 32, 128, 4), (1, 3, 5, 300, 1)])
def test_qkv_epilogue_and_joint_attention(B, h, wd, T, H, gemm_mode, fmha_mode):
    g = gen(8)
    S_img, D = h * wd, H * 128
    S = S_img + T
    w = dict(synthetic.synthetic_weights(1, seed=5, num_heads=H, joint_dim=256, norm_jitter=0.1))
    p = "transformer_blocks.0."
    img, txt = torch.randn(B, S_img, D, generator=g).bfloat16(), torch.randn(B, T, D, generator=g).bfloat16()
    rope = O.rope_tables(1, h, wd, T)
    jq, jk, jv = _qkv_reference(w, p, img, txt, H, rope)
    qd = torch.zeros(B, H, S, 128, dtype=bf, device=dev)
    kd, vd = torch.zeros_like(qd), torch.zeros_like(qd)
    wdv = {k: v.to(dev) for k, v in w.items() if k.startswith(p + "attn")}
    ropd = [t.to(bf).contiguous().to(dev) for t in rope]
    imgd, txtd = img.view(-1, D).to(dev), txt.view(-1, D).to(dev)
    common = dict(N=3 * D, K=D, q=qd.data_ptr(), k=kd.data_ptr(), v=vd.data_ptr(), S_joint=S, H=H, eps=1e-6)
    p0 = q.GemmProblem(A=imgd.data_ptr(), W=wdv[p + "attn.to_qkv.weight"].data_ptr(), bias=wdv[p + "attn.to_qkv.bias"].data_ptr(),
                       M=B * S_img, rows_per_batch=S_img, norm_q_w=wdv[p + "attn.norm_q.weight"].data_ptr(),
                       norm_k_w=wdv[p + "attn.norm_k.weight"].data_ptr(), rope_cos=ropd[0].data_ptr(), rope_sin=ropd[1].data_ptr(),
                       pos_off=T, **common)
    p1 = q.GemmProblem(A=txtd.data_ptr(), W=wdv[p + "attn.add_kv_proj.weight"].data_ptr(), bias=wdv[p + "attn.add_kv_proj.bias"].data_ptr(),
                       M=B * T, rows_per_batch=T, norm_q_w=wdv[p + "attn.norm_added_q.weight"].data_ptr(),
                       norm_k_w=wdv[p + "attn.norm_added_k.weight"].data_ptr(), rope_cos=ropd[2].data_ptr(), rope_sin=ropd[3].data_ptr(),
                       pos_off=0, **common)
    q.gemm([p0, p1], q.EPI_QKV)
    for got, ref in ((qd, jq), (kd, jk), (vd, jv)):
        assert O.rel_fro(got.permute(0, 2, 1, 3).cpu(), ref) < 2e-3
    # attention on the oracle's q/k/v isolates the FMHA kernel
    o_ref = O.joint_attention(jq.float(), jk.float(), jv.float(), 128 ** -0.5).flatten(2, 3)
    qo, ko, vo = (t.permute(0, 2, 1, 3).contiguous().to(dev) for t in (jq, jk, jv))
    ot, oi = q.fmha_joint(qo, ko, vo, T, 128 ** -0.5)
    assert O.rel_fro(ot.cpu().view(B, T, D), o_ref[:, :T]) < TOL_KERNEL
    assert O.rel_fro(oi.cpu().view(B, S_img, D), o_ref[:, T:]) < TOL_KERNEL


def test_attention_backend_plugin_matches_sdpa(fmha_mode):
    from vllm_omni_b200.diffusion.attention.layer import Attention
    g = gen(9)
    qq, kk, vv = (torch.randn(2, 300, 3, 128, generator=g).bfloat16() for _ in range(3))
    attn = Attention(num_heads=3, head_size=128, causal=False, softmax_scale=128 ** -0.5)
    out = attn(qq.to(dev), kk.to(dev), vv.to(dev)).cpu()
    ref = O.joint_attention(qq.float(), kk.float(), vv.float(), 128 ** -0.5)
    assert out.shape == ref.shape and O.rel_fro(out, ref) < TOL_KERNEL


def test_fmha_large_scores_lazy_rescale(fmha_mode):
    """Growing score magnitudes along kv force the lazy O-rescale path (threshold 2^8); tile-to-tile jumps stay far
    below the fast pipeline's 2^100 range, so both pipelines must be exact and the guard must stay silent."""
    g = gen(10)
    B, H, S = 1, 1, 1024
    qq = torch.randn(B, S, H, 128, generator=g).bfloat16() * 4
    kk = torch.randn(B, S, H, 128, generator=g) * torch.linspace(0.1, 6.0, S).view(1, S, 1, 1)
    kk, vv = kk.bfloat16(), torch.randn(B, S, H, 128, generator=g).bfloat16()
    ref = O.joint_attention(qq.float(), kk.float(), vv.float(), 128 ** -0.5).flatten(2, 3)
    qo, ko, vo = (t.permute(0, 2, 1, 3).contiguous().to(dev) for t in (qq, kk, vv))
    q.fmha_overflow(reset=True)
    _, oi = q.fmha_joint(qo, ko, vo, 0, 128 ** -0.5)
    got = oi.cpu().view(B, S, 128)
    assert not q.fmha_overflow()
    assert not torch.isnan(got).any() and O.rel_fro(got, ref) < TOL_KERNEL


def _adversarial_qkv(S, spike_pos, spike_nats, seed):
    """Unit-scale scores everywhere except the keys at `spike_pos`, whose score against EVERY query is `spike_nats`
    nats above the rest: q has a constant component along e_0, the spike keys carry the matching weight there."""
    g = gen(seed)
    qq = torch.randn(1, S, 1, 128, generator=g) * 0.5
    kk = torch.randn(1, S, 1, 128, generator=g) * 0.5
    vv = torch.randn(1, S, 1, 128, generator=g)
    qq[..., 0], kk[..., 0] = 8.0, 0.0
    for p in spike_pos:
        kk[0, p, 0, 0] = spike_nats * (128 ** 0.5) / 8.0   # (q . k) / sqrt(128) gains spike_nats
    return qq.bfloat16(), kk.bfloat16(), vv.bfloat16()


@pytest.mark.parametrize("spike_pos,nats", [([1024 - 3], 120.0), ([517], 120.0), ([300, 900], 90.0), ([1024 - 3], 40.0)],
                         ids=["last-tile+120", "mid-tile+120", "two-spikes+90", "last-tile+40-in-range"])
def test_fmha_adversarial_score_jump_guard(spike_pos, nats):
    """The guard of the default attention pipeline (weak #2 of the round-1 verdict).  A key whose score sits `nats` above
    everything before it, in the LAST KV tile or in the MIDDLE of one: the exact pipeline must match fp32 SDPA to 2^-7;
    the fast pipeline must either match too (jump inside its 2^100 = 69-nat range) or RAISE ITS FLAG — it may never
    return a silently wrong result.  (Round 1's kernel clamped at 2^96 without telling anyone: it fails this test.)"""
    S = 1024
    qq, kk, vv = _adversarial_qkv(S, spike_pos, nats, seed=40)
    ref = O.joint_attention(qq.float(), kk.float(), vv.float(), 128 ** -0.5).flatten(2, 3)
    qo, ko, vo = (t.permute(0, 2, 1, 3).contiguous().to(dev) for t in (qq, kk, vv))
    _, oe = q.fmha_joint(qo, ko, vo, 0, 128 ** -0.5, mode=q.FMHA_EXACT)
    assert O.rel_fro(oe.cpu().view(1, S, 128), ref) < TOL_KERNEL
    for fast in (q.FMHA_FAST, q.FMHA_FAST | 8):          # the fast pipeline, without and with the polynomial share
        q.fmha_overflow(reset=True)
        _, of = q.fmha_joint(qo, ko, vo, 0, 128 ** -0.5, mode=fast)
        flagged = q.fmha_overflow(reset=True)
        err = O.rel_fro(of.cpu().view(1, S, 128).nan_to_num(1e30), ref)
        print(f"spike {spike_pos} +{nats} nats: mode {fast} flagged={flagged} err={err:.3e}")
        assert flagged == (nats * 1.4427 > 100.0)          # 69.3 nats = 2^100
        assert flagged or err < TOL_KERNEL
        assert not q.fmha_overflow()                         # reset worked


def test_denoise_falls_back_to_exact_attention_when_flagged():
    """Pipeline-level guard: norm_q / norm_k weights of ~10 give score standard deviations of ~100 nats, the fast
    pipeline flags the first denoise, the pipeline switches the process to the exact pipeline and recomputes: the result
    is bit-identical to a denoise that ran the exact pipeline from the start (whose kernel is checked against fp32 SDPA
    above; with near-one-hot attention an oracle comparison would measure arg-max ties, not the kernel)."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 1, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}))
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    w = dict(synthetic.synthetic_weights(L, seed=41, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    for k in w:
        if k.endswith(("attn.norm_q.weight", "attn.norm_k.weight")):
            w[k] = (w[k].float() * 10).bfloat16()
    pipe.transformer.load_weights(w.items())
    g = gen(42)
    B, hh, ww, T = 1, 16, 16, 40   # S = 296: three KV tiles
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    pe = torch.randn(B, T, joint, generator=g).bfloat16()
    req = OmniDiffusionRequest(prompt_embeds=pe, latents=lat, height=hh * 16, width=ww * 16, num_inference_steps=2,
                               true_cfg_scale=1.0, output_type="latent")
    prev = q.get_fmha_mode()
    try:
        q.set_fmha_mode(q.FMHA_EXACT)
        want = pipe.forward(req).output.clone()
        q.set_fmha_mode(q.FMHA_FAST)
        out = pipe.forward(req)
        assert out.error is None
        assert q.get_fmha_mode() & 7 == q.FMHA_EXACT, "the guard did not fire on 10x norm weights"
        assert not torch.isnan(out.output).any() and torch.equal(out.output, want)
    finally:
        q.set_fmha_mode(prev)


@pytest.mark.parametrize("name", ["tiny_L2_H2", "narrow_L1_H4_ragged", "fullwidth_L1", "tiny_edit_two_grids", "tiny_edit_three_grids"])
def test_model_forward_vs_reference_golden(golden_dir, name, gemm_mode, fmha_mode):
    fx = torch.load(os.path.join(golden_dir, name + ".pt"))
    c = fx["case"]
    m = make_model(c["L"], c["H"], c["joint"], c["seed"])
    h, w_ = c["grid"]
    shapes = [[(1, h, w_)] + [(1, a, b) for a, b in c.get("extra_grids", [])]] * c["B"]  # edit layout: condition image appended
    args = (fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None)
    out = m(*args, fx["timestep"].to(dev), shapes, [c["T"]] * c["B"], return_dict=False)[0].cpu()
    e_ref = O.rel_fro(out, fx["ref_bf16"])
    e_fp32 = O.rel_fro(out, fx["ref_fp32"])
    print(f"{name}: native vs ref-bf16 {e_ref:.3e}; native vs fp32 {e_fp32:.3e}; ref-bf16 vs fp32 {fx['ref_bf16_vs_fp32']:.3e}")
    assert e_ref <= 1e-2                                   # criterion (ii)
    assert e_fp32 <= fx["ref_bf16_vs_fp32"] + 1e-2         # criterion (iii)
    out_u = m(*args, fx["timestep"][:1].to(dev), shapes, [c["T"]] * c["B"], return_dict=False,
              uniform_timestep=True)[0].cpu()
    assert torch.equal(out_u, out)                          # shared-timestep fast path is exact


def test_model_depth12_fullwidth_criterion_iii(golden_dir):
    """12 full-width blocks (D=3072, H=24): the reference's bf16 path is itself 1.26e-2 from its fp32 path here, so
    the bar is err(native, fp32) <= err(reference-bf16, fp32) + 1e-2; both numbers are printed."""
    fx = torch.load(os.path.join(golden_dir, "fullwidth_L12.pt"))
    c = fx["case"]
    m = make_model(c["L"], c["H"], c["joint"], c["seed"])
    h, w_ = c["grid"]
    out = m(fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"].to(dev),
            [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"], return_dict=False)[0].cpu()
    e_ref, e_fp32 = O.rel_fro(out, fx["ref_bf16"]), O.rel_fro(out, fx["ref_fp32"])
    print(f"L12 full width: native vs ref-bf16 {e_ref:.3e}; native vs fp32 {e_fp32:.3e}; ref-bf16 vs fp32 {fx['ref_bf16_vs_fp32']:.3e}")
    assert not torch.isnan(out).any()
    assert e_fp32 <= fx["ref_bf16_vs_fp32"] + 1e-2
    assert e_ref <= 2.5e-2  # two bf16 evaluation orders at depth 12; each is ~1.2e-2 from fp32


def test_headline_shape_vs_reference_golden(golden_dir, gemm_mode):
    """BASELINE configs[1]'s shape — 1024 px: 64x64 latent grid, T=128, D=3072, H=24, S=4224 (33 KV tiles, 17 query-tile
    pairs, the full GEMM raster) — against the UNMODIFIED reference's bf16 and fp32 outputs (fullwidth_1024px_L2.pt).
    Criterion (ii): <= 1e-2 vs reference-bf16."""
    fx = torch.load(os.path.join(golden_dir, "fullwidth_1024px_L2.pt"))
    c = fx["case"]
    m = make_model(c["L"], c["H"], c["joint"], c["seed"])
    h, w_ = c["grid"]
    q.fmha_overflow(reset=True)
    out = m(fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"].to(dev),
            [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"], return_dict=False)[0].cpu()
    e_ref, e_fp32 = O.rel_fro(out, fx["ref_bf16"]), O.rel_fro(out, fx["ref_fp32"])
    print(f"1024px L2 (S=4224, D=3072): native vs ref-bf16 {e_ref:.3e}; native vs fp32 {e_fp32:.3e}; "
          f"ref-bf16 vs fp32 {fx['ref_bf16_vs_fp32']:.3e}")
    assert not q.fmha_overflow()
    assert e_ref <= 1e-2
    assert e_fp32 <= fx["ref_bf16_vs_fp32"] + 1e-2


def test_full_depth_L60_criterion_iii(golden_dir):
    """FULL DEPTH, 60 blocks (H=8, D=1024, 16x16 grid; the unmodified reference ran in bf16 AND fp32 on CPU):
    err(native, fp32) <= err(reference-bf16, fp32) + 1e-2, both numbers printed (SURVEY §8d criterion (iii))."""
    fx = torch.load(os.path.join(golden_dir, "narrow_L60_H8.pt"))
    c = fx["case"]
    m = make_model(c["L"], c["H"], c["joint"], c["seed"])
    h, w_ = c["grid"]
    out = m(fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"].to(dev),
            [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"], return_dict=False)[0].cpu()
    e_ref, e_fp32 = O.rel_fro(out, fx["ref_bf16"]), O.rel_fro(out, fx["ref_fp32"])
    print(f"L60 (H=8): native vs fp32 {e_fp32:.3e}; reference-bf16 vs fp32 {fx['ref_bf16_vs_fp32']:.3e}; "
          f"native vs ref-bf16 {e_ref:.3e}")
    assert not torch.isnan(out).any()
    assert e_fp32 <= fx["ref_bf16_vs_fp32"] + 1e-2
    assert e_ref <= 2 * fx["ref_bf16_vs_fp32"] + 1e-2   # two bf16 evaluation orders, each that far from fp32


def test_bench_batch_block_vs_oracle():
    """One full-width block at the BENCH batch (B=4, 1024 px, T=128: M_img=16384, M_txt=512, 96 heads x 17 pairs)
    against the CPU oracle's bf16 path, which the headline golden pins to the reference at this shape."""
    L, H, joint = 1, 24, 3584
    m = make_model(L, H, joint, seed=21, norm_jitter=0.1)
    w = dict(synthetic.synthetic_weights(L, seed=21, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    lat, txt = synthetic.synthetic_inputs(4, 1024, 1024, 128)
    t = torch.tensor([0.731]).bfloat16()
    ref = O.model_forward(w, O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint), lat, txt, t.expand(4), (1, 64, 64))
    out = m(lat.to(dev), txt.to(dev), None, t.to(dev), [[(1, 64, 64)]] * 4, [128] * 4, return_dict=False,
            uniform_timestep=True)[0].cpu()
    e = O.rel_fro(out, ref)
    print(f"B=4 1024px one block: native vs oracle-bf16 {e:.3e}")
    assert e <= 1e-2


def test_block_outputs_vs_oracle_intermediates():
    """Residual streams after the last block (img and txt) against the oracle's bf16 path."""
    L, H, joint = 2, 2, 256
    m = make_model(L, H, joint, seed=11)
    w = dict(synthetic.synthetic_weights(L, seed=11, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    g = gen(12)
    hs, eh = torch.randn(2, 48, 64, generator=g).bfloat16(), torch.randn(2, 24, joint, generator=g).bfloat16()
    t = torch.tensor([0.4, 0.4]).bfloat16()
    dims = O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint)
    out_ref, inter = O.model_forward(w, dims, hs, eh, t, (1, 8, 6), return_intermediates=True)
    out = m(hs.to(dev), eh.to(dev), None, t.to(dev), [[(1, 8, 6)]] * 2, [24, 24], return_dict=False)[0]
    img, txt = m.debug_streams(2, 48, 24, torch.device(dev, 0))
    assert O.rel_fro(img.cpu(), inter[-1][1]) < 1e-2 and O.rel_fro(txt.cpu(), inter[-1][0]) < 1e-2
    assert O.rel_fro(out.cpu(), out_ref) < 1e-2


@pytest.mark.parametrize("cfg", [False, True])
def test_diffuse_trajectory_vs_oracle(cfg):
    """4-step denoise through the pipeline's diffuse() (BASELINE config 1 shape class: 4 steps, B=1..2)."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}))
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    w = dict(synthetic.synthetic_weights(L, seed=13, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    pipe.transformer.load_weights(w.items())
    g = gen(14)
    B, hh, ww, T = 2, 8, 6, 20
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    pe, ne = torch.randn(B, T, joint, generator=g).bfloat16(), torch.randn(B, T, joint, generator=g).bfloat16()
    sig = O.flow_match_sigmas(4, hh * ww)
    ref = O.diffuse(w, O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint), lat, pe, ne if cfg else None, sig, (1, hh, ww), 4.0)
    req = OmniDiffusionRequest(prompt_embeds=pe, negative_prompt_embeds=ne if cfg else None, latents=lat, height=hh * 16,
                               width=ww * 16, num_inference_steps=4, true_cfg_scale=4.0 if cfg else 1.0, output_type="latent")
    out = pipe.forward(req)
    assert out.error is None and out.output.shape == lat.shape
    assert np.array_equal(pipe.scheduler.sigmas.numpy(), sig)
    assert O.rel_fro(out.output.cpu(), ref) < 1e-2


@pytest.mark.parametrize("cfg", [False, True])
def test_edit_diffuse_trajectory_vs_oracle(cfg):
    """Image-edit layout (reference pipeline_qwen_image_edit.py:574-639): condition latents appended on the sequence
    axis, two RoPE grids, noisy rows of the prediction kept; 4-step trajectory against the oracle."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image_edit import QwenImageEditPipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}), model_class_name="QwenImageEditPipeline")
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImageEditPipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    w = dict(synthetic.synthetic_weights(L, seed=15, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    pipe.transformer.load_weights(w.items())
    g = gen(16)
    B, hh, ww, T, h2, w2 = 2, 8, 6, 20, 4, 6
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    il = torch.randn(B, h2 * w2, 64, generator=g).bfloat16()
    pe, ne = torch.randn(B, T, joint, generator=g).bfloat16(), torch.randn(B, T, joint, generator=g).bfloat16()
    sig = O.flow_match_sigmas(4, hh * ww)
    ref = O.diffuse(w, O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint), lat, pe, ne if cfg else None, sig,
                    [(1, hh, ww), (1, h2, w2)], 4.0, image_latents=il)
    req = OmniDiffusionRequest(prompt_embeds=pe, negative_prompt_embeds=ne if cfg else None, latents=lat, height=hh * 16,
                               width=ww * 16, num_inference_steps=4, true_cfg_scale=4.0 if cfg else 1.0, output_type="latent",
                               extra={"image_latents": il, "image_latent_grid": (h2, w2)})
    out = pipe.forward(req)
    assert out.error is None and out.output.shape == lat.shape
    assert O.rel_fro(out.output.cpu(), ref) < 1e-2
    # without a condition image the edit pipeline is the text-to-image path
    req2 = OmniDiffusionRequest(prompt_embeds=pe, latents=lat, height=hh * 16, width=ww * 16, num_inference_steps=2,
                                true_cfg_scale=1.0, output_type="latent")
    ref2 = O.diffuse(w, O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint), lat, pe, None, O.flow_match_sigmas(2, hh * ww), (1, hh, ww))
    assert O.rel_fro(pipe.forward(req2).output.cpu(), ref2) < 1e-2


@pytest.mark.parametrize("cfg", [False, True])
def test_cuda_graph_replay_bit_identical_to_eager_loop(cfg):
    """`enable_cuda_graph`: one captured denoise timestep (forward(s) + fused CFG / Euler kernel with the timestep and
    sigmas read from device memory) replayed per step == the eager loop, bit for bit, also on a second request that
    reuses the captured graph with other inputs."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}))
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    pipe.transformer.load_weights(synthetic.synthetic_weights(L, seed=23, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    g = gen(24)

    def mk():
        return OmniDiffusionRequest(prompt_embeds=torch.randn(2, 20, joint, generator=g).bfloat16(),
                                    negative_prompt_embeds=torch.randn(2, 17, joint, generator=g).bfloat16() if cfg else None,
                                    latents=torch.randn(2, 48, 64, generator=g).bfloat16(), height=128, width=96,
                                    num_inference_steps=5, true_cfg_scale=4.0 if cfg else 1.0, output_type="latent")
    r1, r2 = mk(), mk()
    eager = [pipe.forward(r).output.clone() for r in (r1, r2)]
    pipe.enable_cuda_graph(True)
    n0 = q.launch_count()
    graph = [pipe.forward(r).output.clone() for r in (r1, r2)]
    assert len(pipe._graphs) == 1                       # one bucket, captured once, replayed for both requests
    assert q.launch_count() - n0 < 200                  # warm-up + capture only: replays launch nothing through the C ABI
    assert torch.equal(graph[0], eager[0]) and torch.equal(graph[1], eager[1])
    pipe.enable_cuda_graph(False)


def test_edit_plus_two_condition_images_vs_oracle():
    """Edit-plus layout (reference pipeline_qwen_image_edit_plus.py:436-464,729-737): two condition images of different
    sizes appended after the noisy latents, three RoPE grids; 3-step true-CFG trajectory against the oracle."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image_edit import QwenImageEditPlusPipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}), model_class_name="QwenImageEditPlusPipeline")
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImageEditPlusPipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    w = dict(synthetic.synthetic_weights(L, seed=17, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    pipe.transformer.load_weights(w.items())
    g = gen(18)
    B, hh, ww, T = 2, 8, 6, 20
    grids2 = [(4, 6), (3, 5)]
    lat = torch.randn(B, hh * ww, 64, generator=g).bfloat16()
    ils = [torch.randn(B, a * b, 64, generator=g).bfloat16() for a, b in grids2]
    pe, ne = torch.randn(B, T, joint, generator=g).bfloat16(), torch.randn(B, T, joint, generator=g).bfloat16()
    sig = O.flow_match_sigmas(3, hh * ww)
    ref = O.diffuse(w, O.DiTDims(num_layers=L, num_heads=H, joint_dim=joint), lat, pe, ne, sig,
                    [(1, hh, ww)] + [(1, a, b) for a, b in grids2], 4.0, image_latents=torch.cat(ils, dim=1))
    req = OmniDiffusionRequest(prompt_embeds=pe, negative_prompt_embeds=ne, latents=lat, height=hh * 16, width=ww * 16,
                               num_inference_steps=3, true_cfg_scale=4.0, output_type="latent",
                               extra={"image_latents": ils, "image_latent_grids": grids2})
    out = pipe.forward(req)
    assert out.error is None and out.output.shape == lat.shape
    assert O.rel_fro(out.output.cpu(), ref) < 1e-2


def test_cross_request_batching_matches_solo_runs():
    """`GPUWorker.execute_model([r0, r1, r2])`: r0 and r2 (same geometry / schedule / text length) share one denoise batch,
    r1 (another text length) runs alone; every request's latents are bit-identical to running it by itself."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    from vllm_omni_b200.diffusion.worker.gpu_worker import GPUWorker, merge_requests
    L, H, joint = 2, 2, 256
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": L}), synthetic_weights_seed=3)
    wk = GPUWorker.__new__(GPUWorker)
    wk.local_rank, wk.rank, wk.od_config, wk.cache_backend = 0, 0, od, None
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            wk.pipeline = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=H, joint_attention_dim=joint))
    finally:
        torch.set_default_dtype(torch.float32)
    wk.pipeline.transformer.load_weights(synthetic.synthetic_weights(L, seed=3, norm_jitter=0.1, num_heads=H, joint_dim=joint))
    g = gen(19)

    def mk(n, T, seed):
        return OmniDiffusionRequest(prompt_embeds=torch.randn(n, T, joint, generator=g).bfloat16(), height=128, width=96,
                                    num_inference_steps=3, true_cfg_scale=1.0, output_type="latent", seed=seed)
    reqs = [mk(1, 20, 7), mk(2, 33, 8), mk(2, 20, 9)]
    groups = merge_requests(reqs)
    assert [[i for i, _ in parts] for _, parts in groups] == [[0, 2], [1]]
    out = wk.execute_model(reqs, od)
    assert out.error is None and out.trajectory_timesteps == [1, 2, 2] and out.output.shape[0] == 5
    solo = torch.cat([wk.execute_model([r], od).output for r in reqs], dim=0)
    assert torch.equal(out.output, solo)


def test_step_cache_kernels_bit_exact():
    """rel-L1 sums, residual subtraction and residual add (reference cache/teacache/hook.py:131,152,198-203)."""
    g = gen(30)
    a = torch.randn(3, 77, 256, generator=g).bfloat16().to(dev)
    b = (a.float() * 1.03 + 0.01 * torch.randn(3, 77, 256, generator=g).to(dev)).bfloat16()
    sums = torch.zeros(2, dtype=torch.float32, device=dev)
    q.rel_l1_sums(a, b, sums)
    want0, want1 = (a - b).abs().float().sum().item(), b.abs().float().sum().item()
    assert abs(sums[0].item() - want0) <= 1e-5 * want0 and abs(sums[1].item() - want1) <= 1e-5 * want1
    out = torch.empty_like(a)
    q.bf16_sub(out, a, b)
    assert torch.equal(out, a - b)
    x = a.clone()
    q.bf16_add_inplace(x, b)
    assert torch.equal(x, a + b)


def test_staged_forward_equals_single_call(golden_dir):
    """PRE -> BLOCKS -> POST through qimg_engine_forward_stages == qimg_engine_forward, bit for bit (the TeaCache
    hook with a threshold that never reuses runs exactly this sequence)."""
    from vllm_omni_b200.diffusion.cache.teacache import TeaCacheConfig, apply_teacache_hook
    fx = torch.load(os.path.join(golden_dir, "tiny_L2_H2.pt"))
    c = fx["case"]
    m = make_model(c["L"], c["H"], c["joint"], c["seed"])
    h, w_ = c["grid"]
    args = (fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"].to(dev),
            [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"])
    ref = m(*args, return_dict=False)[0].clone()
    # constant polynomial 1.0 >= threshold: every step recomputes, i.e. the hook runs exactly PRE -> BLOCKS -> POST
    apply_teacache_hook(m, TeaCacheConfig(rel_l1_thresh=0.5, coefficients=[0.0, 0.0, 0.0, 0.0, 1.0]))
    for _ in range(3):
        assert torch.equal(m(*args, return_dict=False)[0], ref)
    assert [d[1] for d in m._teacache.decisions] == [True, True, True]


@pytest.mark.parametrize("cfg", [False, True])
def test_teacache_trajectory_vs_reference_hook(golden_dir, cfg):
    """8-step denoise with the native TeaCache hook against the fixture produced by the UNMODIFIED reference hook
    (tests/golden/teacache_tiny.pt): the same compute / reuse decisions (separate positive / negative states under CFG)
    and latents within 1e-2."""
    from vllm_omni_b200.diffusion.cache import get_cache_backend
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    fx = torch.load(os.path.join(golden_dir, "teacache_tiny.pt"), weights_only=False)
    c = fx["case"]
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": c["L"]}), cache_backend="tea_cache",
                             cache_config={"rel_l1_thresh": c["thresh"], "coefficients": c["coefficients"]})
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=c["H"], joint_attention_dim=c["joint"]))
    finally:
        torch.set_default_dtype(torch.float32)
    pipe.transformer.load_weights(synthetic.synthetic_weights(c["L"], seed=c["seed"], norm_jitter=0.1, num_heads=c["H"],
                                                              joint_dim=c["joint"]))
    backend = get_cache_backend(od.cache_backend, od.cache_config)
    backend.enable(pipe)
    h, w_ = c["grid"]
    want = fx["cfg" if cfg else "nocfg"]
    for _ in range(2):  # the second run checks refresh()
        backend.refresh(pipe, c["steps"])
        req = OmniDiffusionRequest(prompt_embeds=fx["prompt_embeds"], negative_prompt_embeds=fx["negative_prompt_embeds"] if cfg else None,
                                   latents=fx["latents0"], height=h * 16, width=w_ * 16, num_inference_steps=c["steps"],
                                   sigmas=None, true_cfg_scale=4.0 if cfg else 1.0, output_type="latent")
        out = pipe.forward(req)
        assert out.error is None
        assert np.array_equal(pipe.scheduler.sigmas.numpy(), fx["sigmas"])
        got = [d[1] for d in pipe.transformer._teacache.decisions]
        print("decisions", got, "rel", [round(d[2], 4) for d in pipe.transformer._teacache.decisions])
        assert got == want["decisions"]
        assert O.rel_fro(out.output.cpu(), want["latents"]) < 1e-2


@pytest.mark.parametrize("cfg", [False, True])
def test_diffuse_vs_reference_loop_golden(golden_dir, cfg):
    """Native denoise loop (engine forwards + fused CFG / Euler kernel) against the latents the reference's OWN
    `QwenImagePipeline.diffuse` produced on CPU (tests/golden/diffuse_tiny.pt; negative prompt of a different length)."""
    from vllm_omni_b200.diffusion.data import OmniDiffusionConfig, TransformerConfig
    from vllm_omni_b200.diffusion.models.qwen_image.pipeline_qwen_image import QwenImagePipeline
    from vllm_omni_b200.diffusion.request import OmniDiffusionRequest
    fx = torch.load(os.path.join(golden_dir, "diffuse_tiny.pt"), weights_only=False)
    c = fx["case"]
    od = OmniDiffusionConfig(tf_model_config=TransformerConfig.from_dict({"num_layers": c["L"]}))
    torch.set_default_dtype(bf)
    try:
        with torch.device(dev):
            pipe = QwenImagePipeline(od_config=od, transformer_kwargs=dict(num_attention_heads=c["H"], joint_attention_dim=c["joint"]))
    finally:
        torch.set_default_dtype(torch.float32)
    pipe.transformer.load_weights(synthetic.synthetic_weights(c["L"], seed=c["seed"], norm_jitter=0.1, num_heads=c["H"],
                                                              joint_dim=c["joint"]))
    h, w_ = c["grid"]
    req = OmniDiffusionRequest(prompt_embeds=fx["prompt_embeds"], negative_prompt_embeds=fx["negative_prompt_embeds"] if cfg else None,
                               latents=fx["latents0"], height=h * 16, width=w_ * 16, num_inference_steps=c["steps"],
                               true_cfg_scale=c["true_cfg_scale"] if cfg else 1.0, output_type="latent")
    out = pipe.forward(req)
    assert out.error is None and np.array_equal(pipe.scheduler.sigmas.numpy(), fx["sigmas"])
    assert O.rel_fro(out.output.cpu(), fx["cfg" if cfg else "nocfg"]) < 1e-2


def test_full_size_properties_1024px():
    """BASELINE configs[1] sizes (1024px, S_img=4096, T=128, D=3072; depth cut to L=2 to bound memory/time):
    size-independent properties — batch rows are independent and deterministic; CFG with identical branches is the
    identity on the noise prediction; the Euler update is linear in dt."""
    m = make_model(2, 24, 3584, seed=15, norm_jitter=0.0)
    lat, txt = synthetic.synthetic_inputs(2, 1024, 1024, 128)
    lat[1], txt[1] = lat[0], txt[0]
    t = torch.tensor([0.6]).bfloat16().to(dev)
    grid = [[(1, 64, 64)]] * 2
    o2 = m(lat.to(dev), txt.to(dev), None, t, grid, [128, 128], return_dict=False, uniform_timestep=True)[0]
    o1 = m(lat[:1].to(dev), txt[:1].to(dev), None, t, grid[:1], [128], return_dict=False, uniform_timestep=True)[0]
    assert not torch.isnan(o2).any()
    assert torch.equal(o2[0], o2[1]) and torch.equal(o1[0], o2[0])
    x = lat[:1].to(dev).clone()
    q.cfg_euler_step(o1, o1.clone(), x, 4.0, 0.9, 0.8)
    y = lat[:1].to(dev).clone()
    q.cfg_euler_step(o1, None, y, 1.0, 0.9, 0.8)
    assert O.rel_fro(x.cpu(), y.cpu()) < 2e-3
    z = lat[:1].to(dev).clone()
    q.cfg_euler_step(o1, None, z, 1.0, 0.9, 0.9)
    assert torch.equal(z.cpu(), lat[:1])
    # one full-width block against the CPU oracle at reduced S (the oracle at S=4224 is covered by the golden 'fullwidth_L1')


def _torchrun(script, nproc, env=None, timeout=600):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", "29531", os.path.join(root, "tools", script)]
    r = subprocess.run(cmd, cwd=root, env=e, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_worker_matches_single_gpu():
    """Image-sharded data parallelism through the worker protocol: gathered latents bit-equal to one GPU's."""
    _torchrun("dp_check.py", 2)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("comm", ["nccl", "p2p"])
def test_tp2_matches_single_gpu(comm):
    """Head/FFN-sharded tensor parallelism (NCCL all-reduce, and the fused peer-memory reduction kernel) within 1e-2 of
    the unsharded engine (the reference's SP-vs-baseline tolerance, test_ulysses_sequence_parallel.py:332-343)."""
    out = _torchrun("tp_check.py", 2, env={"TP_COMM": comm, "TP_LAYERS": "2", "TP_RES": "512"})
    assert "rel_fro" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sp2_bit_identical_to_single_gpu():
    """Fused sequence parallelism (Ulysses: the two all-to-alls as peer stores of the QKV-GEMM / attention epilogues):
    every dot product runs over its full K on one GPU, so the output must equal the single-GPU engine's BIT FOR BIT —
    at a ragged resolution too (rows and heads split unevenly over tiles) and on the full-depth reference fixture."""
    out = _torchrun("tp_check.py", 2, env={"TP_CASES": "sp,2,512,1;sp,3,272,2"})
    assert out.count("bit-identical: True") == 2, out[-2000:]
    out = _torchrun("tp_check.py", 2, env={"TP_GOLDEN": "narrow_L60_H8", "TP_COMM": "sp"})
    assert "criterion (iii) ok" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tp2_full_depth_criterion_iii():
    """Tensor parallel at FULL DEPTH (60 blocks, reference-generated fixture narrow_L60_H8): err(TP, fp32 reference) <=
    err(reference-bf16, fp32 reference) + 1e-2 — the fp32 partial sums of the fused push GEMM add no error of their own."""
    out = _torchrun("tp_check.py", 2, env={"TP_GOLDEN": "narrow_L60_H8", "TP_COMM": "p2p"})
    assert "criterion (iii) ok" in out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_cfg_parallel_bit_equal_to_sequential():
    """Positive / negative branch on two GPUs + one all-gather per step == the sequential two-forward path, bit for bit."""
    out = _torchrun("cfg_check.py", 2)
    assert "bit-equal to sequential = True" in out
