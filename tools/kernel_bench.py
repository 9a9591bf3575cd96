"""Per-kernel timing at the bench shapes (B=4, 1024px, T=128, D=3072): CUDA events, L2-sized buffers rotated."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vllm_omni_b200 import lib as q

dev = "cuda"
B, S_img, T, H = int(os.environ.get("KB_B", 4)), 4096, 128, 24
D, FF, S = H * 128, 4 * H * 128, 4096 + 128
Mi, Mt = B * S_img, B * T
bf = torch.bfloat16
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: (torch.randn(*s, generator=g, device=dev, dtype=torch.float32) * sc).to(bf)

def timeit(name, fn, flops=None, bytes_=None, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts); med = sorted(ts)[len(ts) // 2]
    extra = ""
    if flops: extra += f"  {flops / ms / 1e9:8.1f} TFLOP/s (best) {flops / med / 1e9:8.1f} (median)"
    if bytes_: extra += f"  {bytes_ / ms / 1e6:8.1f} GB/s"
    print(f"{name:44s} best {ms:8.3f} ms  median {med:8.3f} ms{extra}", flush=True)

def gemm_case(name, N, K, epi):
    ai, at = rn(Mi, K), rn(Mt, K)
    wi, wt = rn(N, K, sc=K ** -0.5), rn(N, K, sc=K ** -0.5)
    bi, bt = rn(N), rn(N)
    keep = [ai, at, wi, wt, bi, bt]
    kw_i = dict(A=ai.data_ptr(), W=wi.data_ptr(), bias=bi.data_ptr(), M=Mi, N=N, K=K, rows_per_batch=S_img)
    kw_t = dict(A=at.data_ptr(), W=wt.data_ptr(), bias=bt.data_ptr(), M=Mt, N=N, K=K, rows_per_batch=T)
    if epi == q.EPI_QKV:
        qq = torch.empty(B, H, S, 128, dtype=bf, device=dev); kk = torch.empty_like(qq); vv = torch.empty_like(qq)
        nw = torch.ones(128, dtype=bf, device=dev); cs = rn(S_img, 64); ct = rn(T, 64)
        keep += [qq, kk, vv, nw, cs, ct]
        ex = dict(q=qq.data_ptr(), k=kk.data_ptr(), v=vv.data_ptr(), norm_q_w=nw.data_ptr(), norm_k_w=nw.data_ptr(), S_joint=S, H=H, eps=1e-6)
        p = [q.GemmProblem(rope_cos=cs.data_ptr(), rope_sin=cs.data_ptr(), pos_off=T, **kw_i, **ex),
             q.GemmProblem(rope_cos=ct.data_ptr(), rope_sin=ct.data_ptr(), pos_off=0, **kw_t, **ex)]
    else:
        oi, ot = torch.zeros(Mi, N, dtype=bf, device=dev), torch.zeros(Mt, N, dtype=bf, device=dev)
        gt = rn(B, N); keep += [oi, ot, gt]
        ex = dict(ldo=N, gate=gt.data_ptr(), gate_stride=N)
        p = [q.GemmProblem(out=oi.data_ptr(), **kw_i, **ex), q.GemmProblem(out=ot.data_ptr(), **kw_t, **ex)]
    fl = 2.0 * (Mi + Mt) * N * K
    timeit(name, lambda: q.gemm(p, epi), flops=fl)
    # image stream alone
    timeit(name + " [img only]", lambda: q.gemm(p[:1], epi), flops=2.0 * Mi * N * K)
    return keep

print("device", torch.cuda.get_device_name(0), "B", B)
ONLY = set(filter(None, os.environ.get("KB_ONLY", "").split(",")))
want = lambda k: (not ONLY) or k in ONLY
# cuBLAS reference point for the same shapes
if want("cublas"):
    for (N, K) in ((3 * D, D), (D, D), (FF, D), (D, FF)):
        a, w = rn(Mi, K), rn(N, K)
        timeit(f"cuBLAS (torch) {Mi}x{N}x{K}", lambda: torch.nn.functional.linear(a, w), flops=2.0 * Mi * N * K)
if want("qkv"): gemm_case("gemm QKV+norm+rope  N=9216 K=3072", 3 * D, D, q.EPI_QKV)
if want("outproj"): gemm_case("gemm out-proj+gate-res N=3072 K=3072", D, D, q.EPI_BIAS_GATE_RES)
if want("mlpup"): gemm_case("gemm MLP up + GELU  N=12288 K=3072", FF, D, q.EPI_BIAS_GELU)
if want("mlpdown"): gemm_case("gemm MLP down+gate-res N=3072 K=12288", D, FF, q.EPI_BIAS_GATE_RES)
if not (want("fmha") or want("ew")): sys.exit(0)
qq, kk, vv = rn(B, H, S, 128), rn(B, H, S, 128), rn(B, H, S, 128)
ot, oi = torch.empty(Mt, D, dtype=bf, device=dev), torch.empty(Mi, D, dtype=bf, device=dev)
timeit("fmha joint S=4224", lambda: q.fmha_joint(qq, kk, vv, T, 128 ** -0.5, ot, oi), flops=4.0 * B * H * S * S * 128)
if not want("ew"): sys.exit(0)
try:
    qs, ks, vs = (t.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3) for t in (qq, kk, vv))
    timeit("torch SDPA (library) same shape", lambda: torch.nn.functional.scaled_dot_product_attention(qq, kk, vv), flops=4.0 * B * H * S * S * 128)
except Exception as e:
    print("sdpa failed", e)
x = rn(Mi, D); mod = rn(B, 3 * D); y = torch.empty_like(x)
timeit("ln_modulate img", lambda: q.ln_modulate(x, mod[:, :D], mod[:, D:2 * D], S_img, 3 * D, out=y), bytes_=2.0 * Mi * D * 2)
L = 60
temb = rn(1, D); Wm = rn(L * 12 * D // 8, D)  # 1/8 of the full modulation weight (1.7 GB)
timeit("linear_small_m mods (1/8 of 13.6 GB) M=1", lambda: q.linear_small_m(temb, Wm, None, True), bytes_=Wm.numel() * 2.0)
lat = rn(B * S_img, 64); pos = rn(B * S_img, 64); neg = rn(B * S_img, 64)
timeit("cfg_euler_step (cfg)", lambda: q.cfg_euler_step(pos, neg, lat, 4.0, 0.9, 0.85), bytes_=4.0 * B * S_img * 64 * 2)
