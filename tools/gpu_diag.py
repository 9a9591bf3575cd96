"""GPU diagnostics (run under gpurun): each check compares one sm_100a kernel with the CPU oracle and
prints error statistics instead of asserting, so one call yields maximum information.
    python tools/gpu_diag.py [check ...]        (no args = all)
"""
from __future__ import annotations

import os
import sys
import time
import traceback

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import qwen_image_oracle as O  # noqa: E402  (diagnostics are test infrastructure)
from vllm_omni_b200 import lib as q  # noqa: E402
from vllm_omni_b200 import synthetic  # noqa: E402

dev = "cuda"


def stats(name, got, ref):
    got = got.float().cpu()
    ref = ref.float().cpu()
    d = (got - ref).abs()
    rel = O.rel_fro(got, ref)
    print(f"  {name}: rel_fro={rel:.3e} max_abs={d.max():.3e} ref_absmean={ref.abs().mean():.3e} "
          f"nan={int(torch.isnan(got).sum())} mismatch>1e-2={(d > 1e-2 * ref.abs().clamp_min(1e-3)).float().mean():.4f}")
    return rel


def check_probe():
    g = torch.Generator().manual_seed(0)
    A = torch.randn(128, 128, generator=g).bfloat16()
    Bm = torch.randn(128, 128, generator=g).bfloat16()  # mode 0: [N,K]; modes 1,2: [K,N]
    for mode in (0, 1, 2):
        D = q.umma_probe(A.to(dev), Bm.to(dev), mode)
        torch.cuda.synchronize()
        ref = A.float() @ (Bm.float().T if mode == 0 else Bm.float())
        r = stats(f"umma_probe mode {mode}", D, ref)
        if r > 1e-3:
            Dc = D.cpu()
            # help diagnose layout errors: which rows/cols are right?
            ok = ((Dc - ref).abs() < 1e-2 * ref.abs().clamp_min(1.0))
            print("    rows ok frac (first 16):", ok.float().mean(1)[:16].tolist())
            print("    cols ok frac (first 16):", ok.float().mean(0)[:16].tolist())
            alt = A.float() @ (Bm.float() if mode == 0 else Bm.float().T)
            stats("    (vs transposed-B hypothesis)", D, alt)


def check_elementwise():
    g = torch.Generator().manual_seed(1)
    for (B, S, D) in ((2, 40, 256), (1, 33, 3072), (3, 7, 512)):
        x = (torch.randn(B, S, D, generator=g) * 2 + 0.3).bfloat16()
        mod = (torch.randn(B, 3 * D, generator=g) * 0.5).bfloat16()
        ref, _ = O.ada_layer_norm(x, mod, 1e-6)
        modd = mod.to(dev)
        y = q.ln_modulate(x.view(-1, D).to(dev), modd[:, :D], modd[:, D:2 * D], S, 3 * D)
        stats(f"ln_modulate B{B} S{S} D{D}", y.view(B, S, D), ref)
    x = torch.randn(37, 3584, generator=g).bfloat16()
    w = (1 + 0.1 * torch.randn(3584, generator=g)).bfloat16()
    stats("rms_norm 3584", q.rms_norm(x.to(dev), w.to(dev)), O.rms_norm(x, w, 1e-6))
    # gate residual
    x = torch.randn(2, 19, 256, generator=g).bfloat16()
    y = torch.randn(2, 19, 256, generator=g).bfloat16()
    gate = torch.randn(2, 256, generator=g).bfloat16()
    ref = x + gate[:, None, :] * y
    xd = x.view(-1, 256).to(dev).clone()
    q.gate_residual(xd, y.view(-1, 256).to(dev), gate.to(dev), 19, 256)
    stats("gate_residual", xd.view(2, 19, 256), ref)
    # small-M linear
    for (M, N, K, act) in ((1, 1536, 256, True), (4, 777 * 8, 3072, True), (3, 512, 256, False), (11, 640, 512, True)):
        xx = torch.randn(M, K, generator=g).bfloat16()
        W = (torch.randn(N, K, generator=g) * 0.05).bfloat16()
        b = torch.randn(N, generator=g).bfloat16()
        xin = torch.nn.functional.silu(xx) if act else xx
        ref = torch.nn.functional.linear(xin, W, b)
        stats(f"linear_small_m M{M} N{N} K{K}", q.linear_small_m(xx.to(dev), W.to(dev), b.to(dev), act), ref)
    t = (torch.tensor([1000.0, 731.0, 20.5, 0.0]).bfloat16() / 1000)
    stats("timestep_sinusoid", q.timestep_sinusoid(t.to(dev)), O.timestep_sinusoid(t.float()).bfloat16())
    # cfg + euler
    pos = torch.randn(2, 50, 64, generator=g).bfloat16()
    neg = torch.randn(2, 50, 64, generator=g).bfloat16()
    lat = torch.randn(2, 50, 64, generator=g).bfloat16()
    sig, sig_n = torch.tensor(0.9, dtype=torch.float32), torch.tensor(0.85, dtype=torch.float32)
    ref = O.euler_step(O.cfg_combine(pos, neg, 4.0), lat, sig, sig_n)
    ld = lat.to(dev).clone()
    q.cfg_euler_step(pos.to(dev), neg.to(dev), ld, 4.0, float(sig), float(sig_n))
    stats("cfg_euler_step (cfg)", ld, ref)
    ref = O.euler_step(pos, lat, sig, sig_n)
    ld = lat.to(dev).clone()
    q.cfg_euler_step(pos.to(dev), None, ld, 1.0, float(sig), float(sig_n))
    stats("cfg_euler_step (no cfg)", ld, ref)


def check_gemm():
    g = torch.Generator().manual_seed(2)
    F = torch.nn.functional
    for (M, N, K) in ((128, 256, 64), (128, 256, 256), (300, 768, 256), (1024, 3072, 3072), (77, 64, 512), (512, 12288, 3072)):
        x = torch.randn(M, K, generator=g).bfloat16()
        W = (torch.randn(N, K, generator=g) * (1.0 / K ** 0.5)).bfloat16()
        b = torch.randn(N, generator=g).bfloat16()
        ref = F.linear(x.float(), W.float(), b.float())
        t0 = time.time()
        y = q.linear(x.to(dev), W.to(dev), b.to(dev))
        torch.cuda.synchronize()
        r = stats(f"gemm bias M{M} N{N} K{K} ({time.time()-t0:.3f}s)", y, ref)
        if r > 2e-2:
            yc = y.float().cpu()
            ok = ((yc - ref).abs() < 3e-2 * ref.abs().clamp_min(0.5))
            print("    row-block ok:", [round(float(ok[i:i + 32].float().mean()), 2) for i in range(0, min(M, 256), 32)])
            print("    col-block ok:", [round(float(ok[:, j:j + 32].float().mean()), 2) for j in range(0, min(N, 512), 32)])
        if N > 64:
            y = q.linear(x.to(dev), W.to(dev), b.to(dev), q.EPI_BIAS_GELU)
            stats("   + gelu", y, F.gelu(F.linear(x, W, b).float(), approximate="tanh"))
    # grouped gate-residual
    Mi, Mt, D, K = 2 * 200, 2 * 24, 256, 1024
    xi, xt = torch.randn(Mi, D, generator=g).bfloat16(), torch.randn(Mt, D, generator=g).bfloat16()
    ai, at = torch.randn(Mi, K, generator=g).bfloat16(), torch.randn(Mt, K, generator=g).bfloat16()
    Wi, Wt = (torch.randn(D, K, generator=g) / 32).bfloat16(), (torch.randn(D, K, generator=g) / 32).bfloat16()
    bi, bt = torch.randn(D, generator=g).bfloat16(), torch.randn(D, generator=g).bfloat16()
    gi, gt = torch.randn(2, D, generator=g).bfloat16(), torch.randn(2, D, generator=g).bfloat16()
    ref_i = xi.view(2, 200, D) + gi[:, None] * F.linear(ai, Wi, bi).view(2, 200, D)
    ref_t = xt.view(2, 24, D) + gt[:, None] * F.linear(at, Wt, bt).view(2, 24, D)
    xid, xtd = xi.to(dev).clone(), xt.to(dev).clone()
    keep = [ai.to(dev), Wi.to(dev), bi.to(dev), gi.to(dev), at.to(dev), Wt.to(dev), bt.to(dev), gt.to(dev)]
    p0 = q.GemmProblem(A=keep[0].data_ptr(), W=keep[1].data_ptr(), bias=keep[2].data_ptr(), M=Mi, N=D, K=K, rows_per_batch=200,
                       out=xid.data_ptr(), ldo=D, gate=keep[3].data_ptr(), gate_stride=D)
    p1 = q.GemmProblem(A=keep[4].data_ptr(), W=keep[5].data_ptr(), bias=keep[6].data_ptr(), M=Mt, N=D, K=K, rows_per_batch=24,
                       out=xtd.data_ptr(), ldo=D, gate=keep[7].data_ptr(), gate_stride=D)
    q.gemm([p0, p1], q.EPI_BIAS_GATE_RES)
    stats("grouped gate_res img", xid.view(2, 200, D), ref_i)
    stats("grouped gate_res txt", xtd.view(2, 24, D), ref_t)


def _qkv_case(g, B, S_img, T, H):
    D = H * 128
    dims = O.DiTDims(num_layers=1, num_heads=H, joint_dim=256)
    w = {k: v for k, v in synthetic.synthetic_weights(1, seed=5, num_heads=H, joint_dim=256, norm_jitter=0.1)}
    img = torch.randn(B, S_img, D, generator=g).bfloat16()
    txt = torch.randn(B, T, D, generator=g).bfloat16()
    return dims, w, img, txt


def check_qkv_fmha():
    g = torch.Generator().manual_seed(3)
    F = torch.nn.functional
    for (B, h, wd, T, H) in ((1, 8, 16, 128, 2), (2, 10, 9, 37, 2), (1, 32, 32, 128, 4)):
        S_img = h * wd
        S = S_img + T
        D = H * 128
        dims, w, img, txt = _qkv_case(g, B, S_img, T, H)
        p = "transformer_blocks.0."
        rope = O.rope_tables(1, h, wd, T)
        # oracle pieces (bf16 path)
        qkv = F.linear(img, w[p + "attn.to_qkv.weight"], w[p + "attn.to_qkv.bias"])
        iq, ik, iv = (t.unflatten(-1, (H, -1)) for t in qkv.chunk(3, dim=-1))
        qkv = F.linear(txt, w[p + "attn.add_kv_proj.weight"], w[p + "attn.add_kv_proj.bias"])
        tq, tk, tv = (t.unflatten(-1, (H, -1)) for t in qkv.chunk(3, dim=-1))
        iq = O.rms_norm(iq, w[p + "attn.norm_q.weight"], 1e-6); ik = O.rms_norm(ik, w[p + "attn.norm_k.weight"], 1e-6)
        tq = O.rms_norm(tq, w[p + "attn.norm_added_q.weight"], 1e-6); tk = O.rms_norm(tk, w[p + "attn.norm_added_k.weight"], 1e-6)
        bf = torch.bfloat16
        iq = O.apply_rope_interleaved(iq, rope[0].to(bf), rope[1].to(bf)); ik = O.apply_rope_interleaved(ik, rope[0].to(bf), rope[1].to(bf))
        tq = O.apply_rope_interleaved(tq, rope[2].to(bf), rope[3].to(bf)); tk = O.apply_rope_interleaved(tk, rope[2].to(bf), rope[3].to(bf))
        jq, jk, jv = torch.cat([tq, iq], 1), torch.cat([tk, ik], 1), torch.cat([tv, iv], 1)  # [B,S,H,hd]
        # device
        qd = torch.zeros(B, H, S, 128, dtype=bf, device=dev); kd = torch.zeros_like(qd); vd = torch.zeros_like(qd)
        wd_ = {k: v.to(dev) for k, v in w.items() if k.startswith(p + "attn")}
        ropd = [t.to(bf).contiguous().to(dev) for t in rope]
        imgd, txtd = img.view(-1, D).to(dev), txt.view(-1, D).to(dev)
        common = dict(N=3 * D, K=D, q=qd.data_ptr(), k=kd.data_ptr(), v=vd.data_ptr(), S_joint=S, H=H, eps=1e-6)
        p0 = q.GemmProblem(A=imgd.data_ptr(), W=wd_[p + "attn.to_qkv.weight"].data_ptr(), bias=wd_[p + "attn.to_qkv.bias"].data_ptr(),
                           M=B * S_img, rows_per_batch=S_img, norm_q_w=wd_[p + "attn.norm_q.weight"].data_ptr(),
                           norm_k_w=wd_[p + "attn.norm_k.weight"].data_ptr(), rope_cos=ropd[0].data_ptr(), rope_sin=ropd[1].data_ptr(),
                           pos_off=T, **common)
        p1 = q.GemmProblem(A=txtd.data_ptr(), W=wd_[p + "attn.add_kv_proj.weight"].data_ptr(), bias=wd_[p + "attn.add_kv_proj.bias"].data_ptr(),
                           M=B * T, rows_per_batch=T, norm_q_w=wd_[p + "attn.norm_added_q.weight"].data_ptr(),
                           norm_k_w=wd_[p + "attn.norm_added_k.weight"].data_ptr(), rope_cos=ropd[2].data_ptr(), rope_sin=ropd[3].data_ptr(),
                           pos_off=0, **common)
        q.gemm([p0, p1], q.EPI_QKV)
        torch.cuda.synchronize()
        tag = f"B{B} S_img{S_img} T{T} H{H}"
        stats(f"qkv epilogue Q {tag}", qd.permute(0, 2, 1, 3), jq)
        stats(f"qkv epilogue K {tag}", kd.permute(0, 2, 1, 3), jk)
        stats(f"qkv epilogue V {tag}", vd.permute(0, 2, 1, 3), jv)
        # attention on the ORACLE q/k/v (isolates the FMHA kernel)
        o_ref = O.joint_attention(jq.float(), jk.float(), jv.float(), 1.0 / 128 ** 0.5).flatten(2, 3)
        qo, ko, vo = (t.permute(0, 2, 1, 3).contiguous().to(dev) for t in (jq, jk, jv))
        t0 = time.time()
        ot, oi = q.fmha_joint(qo, ko, vo, T, 1.0 / 128 ** 0.5)
        torch.cuda.synchronize()
        print(f"  fmha time {time.time()-t0:.3f}s")
        stats(f"fmha txt {tag}", ot.view(B, T, D), o_ref[:, :T])
        stats(f"fmha img {tag}", oi.view(B, S_img, D), o_ref[:, T:])


def check_model():
    from vllm_omni_b200.diffusion.models.qwen_image.qwen_image_transformer import QwenImageTransformer2DModel
    gd = os.path.join(ROOT, "tests", "golden")
    for name in sorted(os.listdir(gd)):
        if not name.endswith(".pt"):
            continue
        fx = torch.load(os.path.join(gd, name))
        c = fx["case"]
        torch.set_default_dtype(torch.bfloat16)
        with torch.device(dev):
            m = QwenImageTransformer2DModel(num_layers=c["L"], num_attention_heads=c["H"], joint_attention_dim=c["joint"])
        torch.set_default_dtype(torch.float32)
        m.load_weights(synthetic.split_qkv_checkpoint_names(
            synthetic.synthetic_weights(c["L"], seed=c["seed"], norm_jitter=0.1, num_heads=c["H"], joint_dim=c["joint"])))
        h, w_ = c["grid"]
        out = m(fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"].to(dev),
                [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"], return_dict=False)[0]
        torch.cuda.synchronize()
        print(f"  [{name}] ref bf16-vs-fp32 = {fx['ref_bf16_vs_fp32']:.3e}")
        stats(f"model {name} vs ref_bf16", out, fx["ref_bf16"])
        stats(f"model {name} vs ref_fp32", out, fx["ref_fp32"])
        out2 = m(fx["hidden_states"].to(dev), fx["encoder_hidden_states"].to(dev), None, fx["timestep"][:1].to(dev),
                 [[(1, h, w_)]] * c["B"], [c["T"]] * c["B"], return_dict=False, uniform_timestep=True)[0]
        stats(f"model {name} uniform-timestep vs per-sample", out2, out)


CHECKS = dict(probe=check_probe, elementwise=check_elementwise, gemm=check_gemm, qkv_fmha=check_qkv_fmha, model=check_model)

if __name__ == "__main__":
    names = sys.argv[1:] or list(CHECKS)
    print("device:", torch.cuda.get_device_name(0), "sms:", q.device_check())
    for n in names:
        print(f"== {n} ==", flush=True)
        try:
            CHECKS[n]()
            torch.cuda.synchronize()
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()
