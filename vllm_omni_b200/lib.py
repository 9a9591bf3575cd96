"""ctypes binding of the C-ABI library `csrc/libqimg_b200.so` (include/qimg_b200.h).

PyTorch is plumbing here: tensors own the device memory, `tensor.data_ptr()` and the
current stream's handle cross the C ABI.  There is NO fallback path: if the library is
missing or a call fails, a RuntimeError is raised (the reference's worker turns that
into DiffusionOutput(error=...), diffusion/worker/gpu_worker.py:267-274).
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libqimg_b200.so")

EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GATE_RES, EPI_QKV, EPI_PARTIAL_F32 = 0, 1, 2, 3, 4

# every symbol declared in include/qimg_b200.h (tests check the .so exports all of them)
EXPORTED_SYMBOLS = [
    "qimg_abi_version", "qimg_last_error", "qimg_device_check", "qimg_launch_count", "qimg_reset_launch_count",
    "qimg_ln_modulate", "qimg_gate_residual", "qimg_rms_norm", "qimg_linear_small_m", "qimg_timestep_sinusoid",
    "qimg_cfg_euler_step", "qimg_gemm", "qimg_fmha_joint", "qimg_engine_create", "qimg_engine_destroy",
    "qimg_engine_workspace_bytes", "qimg_engine_forward", "qimg_engine_ws_offset_img", "qimg_engine_ws_offset_txt",
    "qimg_umma_probe", "qimg_prof_enable", "qimg_prof_collect", "qimg_set_gemm_mode", "qimg_get_gemm_mode", "qimg_set_fmha_mode", "qimg_get_fmha_mode", "qimg_gate_residual_bias", "qimg_engine_set_tp",
    "qimg_p2p_alloc", "qimg_p2p_free", "qimg_ipc_get_handle", "qimg_ipc_open_handle", "qimg_ipc_close_handle",
    "qimg_engine_set_tp_p2p", "qimg_engine_p2p_error", "qimg_set_fmha_trace",
    "qimg_engine_forward_stages", "qimg_engine_ws_offset_mod", "qimg_rel_l1_sums", "qimg_bf16_sub", "qimg_bf16_add_inplace",
    "qimg_fmha_joint_mode", "qimg_fmha_overflow", "qimg_cfg_euler_step_dev", "qimg_set_euler_dt_fp32",
    "qimg_set_nvtx", "qimg_set_gemm_group_m", "qimg_set_fmha_single_tile", "qimg_ln_modulate_rows", "qimg_fmha_joint_sp",
    "qimg_engine_set_sp_p2p", "qimg_tea_decide", "qimg_tea_residual", "qimg_engine_set_blocks_predicate",
    "qimg_conv2d_nhwc_tf32", "qimg_vae_rms_act", "qimg_vae_upsample2x", "qimg_vae_post_quant", "qimg_vae_conv_out",
    "qimg_vae_softmax_rows", "qimg_vae_transpose", "qimg_set_vae_conv_variant",
    "qimg_conv2d_down2_nhwc_tf32", "qimg_vae_image_to_nhwc", "qimg_ln_modulate_indexed", "qimg_select_rows",
]


class GemmProblem(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("W", C.c_void_p), ("bias", C.c_void_p),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int), ("rows_per_batch", C.c_int),
        ("out", C.c_void_p), ("ldo", C.c_int), ("gate", C.c_void_p), ("gate_stride", C.c_longlong),
        ("q", C.c_void_p), ("k", C.c_void_p), ("v", C.c_void_p),
        ("norm_q_w", C.c_void_p), ("norm_k_w", C.c_void_p), ("rope_cos", C.c_void_p), ("rope_sin", C.c_void_p),
        ("S_joint", C.c_int), ("pos_off", C.c_int), ("H", C.c_int), ("eps", C.c_float),
        ("tp_recv", C.c_void_p * 8), ("tp_size", C.c_int), ("tp_rank", C.c_int), ("tp_recv_rows", C.c_int),
        ("tp_recv_row_off", C.c_int),
        ("row_base", C.c_int), ("sp_size", C.c_int), ("sp_q", C.c_void_p * 8), ("sp_k", C.c_void_p * 8), ("sp_v", C.c_void_p * 8),
    ]


class Dims(C.Structure):
    _fields_ = [("num_layers", C.c_int), ("num_heads", C.c_int), ("head_dim", C.c_int), ("in_channels", C.c_int),
                ("out_dim", C.c_int), ("joint_dim", C.c_int), ("eps", C.c_float)]


BLOCK_FIELDS = [
    "img_mod_w", "img_mod_b", "txt_mod_w", "txt_mod_b", "to_qkv_w", "to_qkv_b", "add_kv_w", "add_kv_b",
    "norm_q", "norm_k", "norm_added_q", "norm_added_k", "to_out_w", "to_out_b", "to_add_out_w", "to_add_out_b",
    "img_mlp_w1", "img_mlp_b1", "img_mlp_w2", "img_mlp_b2", "txt_mlp_w1", "txt_mlp_b1", "txt_mlp_w2", "txt_mlp_b2",
]
GLOBAL_FIELDS = [
    "t_lin1_w", "t_lin1_b", "t_lin2_w", "t_lin2_b", "txt_norm_w", "img_in_w", "img_in_b", "txt_in_w", "txt_in_b",
    "norm_out_w", "norm_out_b", "proj_out_w", "proj_out_b", "mod_all_w", "mod_all_b",
]


# int (*qimg_allreduce_fn)(void* buf, long long count, void* user, qimg_stream_t stream)
ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p)


class BlockWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in BLOCK_FIELDS]


class GlobalWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in GLOBAL_FIELDS]


_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built — never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: the sm_100a CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i, ll, f, sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t
    lib.qimg_abi_version.restype = i
    lib.qimg_last_error.restype = C.c_char_p
    lib.qimg_device_check.argtypes = [C.POINTER(i)]
    lib.qimg_launch_count.restype = ll
    lib.qimg_reset_launch_count.restype = None
    lib.qimg_ln_modulate.argtypes = [vp, vp, vp, vp, i, i, i, ll, f, vp]
    lib.qimg_ln_modulate_rows.argtypes = [vp, vp, vp, vp, i, i, i, i, ll, f, vp]
    lib.qimg_ln_modulate_indexed.argtypes = [vp, vp, vp, vp, i, i, i, ll, f, vp, i, vp]
    lib.qimg_select_rows.argtypes = [vp, ll, vp, vp, i, i, i, i, vp]
    lib.qimg_gate_residual.argtypes = [vp, vp, vp, i, i, i, ll, vp]
    lib.qimg_rms_norm.argtypes = [vp, vp, vp, i, i, f, vp]
    lib.qimg_gate_residual_bias.argtypes = [vp, vp, vp, vp, i, i, i, ll, vp]
    lib.qimg_engine_set_tp.argtypes = [vp, i, ALLREDUCE_FN, vp]
    lib.qimg_p2p_alloc.argtypes = [sz, C.POINTER(vp)]
    lib.qimg_p2p_free.argtypes = [vp]
    lib.qimg_ipc_get_handle.argtypes = [vp, C.c_char_p]
    lib.qimg_ipc_open_handle.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.qimg_ipc_close_handle.argtypes = [vp]
    lib.qimg_engine_set_tp_p2p.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(vp)]
    lib.qimg_engine_set_sp_p2p.argtypes = [vp, i, i, C.POINTER(vp), C.POINTER(vp)]
    lib.qimg_engine_p2p_error.argtypes = [vp, C.POINTER(i)]
    lib.qimg_linear_small_m.argtypes = [vp, vp, vp, vp, i, ll, i, ll, i, vp]
    lib.qimg_timestep_sinusoid.argtypes = [vp, vp, i, vp]
    lib.qimg_cfg_euler_step.argtypes = [vp, vp, vp, ll, i, f, f, f, vp]
    lib.qimg_cfg_euler_step_dev.argtypes = [vp, vp, vp, ll, i, f, vp, vp]
    lib.qimg_set_euler_dt_fp32.argtypes = [i]
    lib.qimg_set_nvtx.argtypes = [i]
    lib.qimg_set_nvtx.restype = None
    lib.qimg_gemm.argtypes = [C.POINTER(GemmProblem), i, i, vp]
    lib.qimg_fmha_joint.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, f, vp]
    lib.qimg_fmha_joint_mode.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, f, i, vp]
    lib.qimg_fmha_overflow.argtypes = [C.POINTER(i), i]
    lib.qimg_engine_create.argtypes = [C.POINTER(Dims), C.POINTER(GlobalWeights), C.POINTER(BlockWeights), C.POINTER(vp)]
    lib.qimg_engine_destroy.argtypes = [vp]
    lib.qimg_engine_destroy.restype = None
    lib.qimg_engine_workspace_bytes.argtypes = [vp, i, i, i]
    lib.qimg_engine_workspace_bytes.restype = sz
    lib.qimg_engine_ws_offset_img.argtypes = [vp, i, i, i]
    lib.qimg_engine_ws_offset_img.restype = sz
    lib.qimg_engine_ws_offset_txt.argtypes = [vp, i, i, i]
    lib.qimg_engine_ws_offset_txt.restype = sz
    lib.qimg_engine_forward.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.qimg_engine_forward_stages.argtypes = [vp, i, vp, vp, vp, i, vp, vp, vp, vp, i, i, i, vp, vp, sz, vp]
    lib.qimg_engine_ws_offset_mod.argtypes = [vp, i, i, i]
    lib.qimg_engine_ws_offset_mod.restype = sz
    lib.qimg_rel_l1_sums.argtypes = [vp, vp, ll, vp, vp]
    lib.qimg_bf16_sub.argtypes = [vp, vp, vp, ll, vp]
    lib.qimg_bf16_add_inplace.argtypes = [vp, vp, ll, vp]
    lib.qimg_tea_decide.argtypes = [vp, ll, C.POINTER(C.c_double), C.c_double, vp, vp, vp, i, i, vp]
    lib.qimg_tea_residual.argtypes = [vp, vp, vp, ll, vp, vp]
    lib.qimg_engine_set_blocks_predicate.argtypes = [vp, vp]
    lib.qimg_umma_probe.argtypes = [vp, vp, vp, i, i, i, vp]
    lib.qimg_conv2d_nhwc_tf32.argtypes = [vp, i, vp, i, vp, vp, i, vp, i, i, i, i, i, i, i, vp]
    lib.qimg_vae_rms_act.argtypes = [vp, vp, vp, ll, i, i, vp]
    lib.qimg_set_vae_conv_variant.argtypes = [i]
    lib.qimg_conv2d_down2_nhwc_tf32.argtypes = [vp, i, vp, i, vp, vp, i, i, i, i, i, i, vp]
    lib.qimg_vae_image_to_nhwc.argtypes = [vp, vp, i, i, i, i, vp]
    lib.qimg_vae_upsample2x.argtypes = [vp, vp, i, i, i, i, vp]
    lib.qimg_vae_post_quant.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    lib.qimg_vae_conv_out.argtypes = [vp, vp, vp, vp, vp, i, i, i, i, vp]
    lib.qimg_vae_softmax_rows.argtypes = [vp, i, i, ll, f, vp]
    lib.qimg_vae_transpose.argtypes = [vp, ll, vp, i, i, vp]
    lib.qimg_set_gemm_mode.argtypes = [i]
    lib.qimg_set_gemm_group_m.argtypes = [i]
    lib.qimg_set_fmha_mode.argtypes = [i]
    lib.qimg_set_fmha_single_tile.argtypes = [i]
    lib.qimg_set_fmha_trace.argtypes = [vp]
    lib.qimg_prof_enable.argtypes = [i]
    lib.qimg_prof_enable.restype = None
    lib.qimg_prof_collect.argtypes = [i, C.POINTER(C.c_double), C.POINTER(ll), C.POINTER(C.c_double)]
    for name in EXPORTED_SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("qimg_abi_version",):
            pass
    _lib = lib
    return lib


def check(rc: int, what: str = "qimg"):
    if rc != 0:
        msg = load().qimg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed: {msg}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    if t is None:
        return None
    assert t.is_cuda, "qimg_b200 ops take CUDA tensors only (no CPU fallback)"
    return t.data_ptr()


def _bf16c(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.bfloat16, f"expected bf16, got {t.dtype}"
    assert t.is_contiguous(), "expected a contiguous tensor"
    return t


def device_check() -> int:
    n = C.c_int(0)
    check(load().qimg_device_check(C.byref(n)), "qimg_device_check")
    return n.value


def launch_count() -> int:
    return int(load().qimg_launch_count())


def reset_launch_count():
    load().qimg_reset_launch_count()


def set_gemm_mode(mode: int):
    check(load().qimg_set_gemm_mode(int(mode)), "qimg_set_gemm_mode")


def get_gemm_mode() -> int:
    return int(load().qimg_get_gemm_mode())


def set_fmha_mode(mode: int):
    check(load().qimg_set_fmha_mode(int(mode)), "qimg_set_fmha_mode")


def set_fmha_single_tile(mode: int):
    """-1 auto, 0 query-tile pairs per CTA, 1 one query tile per CTA (qimg_set_fmha_single_tile)."""
    check(load().qimg_set_fmha_single_tile(int(mode)), "qimg_set_fmha_single_tile")


def get_fmha_mode() -> int:
    return int(load().qimg_get_fmha_mode())


FMHA_EXACT, FMHA_FAST = 4, 6


def fmha_overflow(reset: bool = True) -> bool:
    """True if a fast-pipeline attention launch on the current device flagged an out-of-range score since the last reset
    (synchronising 4-byte read, qimg_fmha_overflow)."""
    v = C.c_int(0)
    check(load().qimg_fmha_overflow(C.byref(v), int(reset)), "qimg_fmha_overflow")
    return v.value != 0


def prof_enable(on: bool):
    load().qimg_prof_enable(int(on))


def prof_collect(kind: int) -> dict:
    """kind 0 = tcgen05 GEMM, 1 = FMHA -> {ms, launches, flops} since the previous collect."""
    ms, n, fl = C.c_double(0), C.c_longlong(0), C.c_double(0)
    check(load().qimg_prof_collect(kind, C.byref(ms), C.byref(n), C.byref(fl)), "qimg_prof_collect")
    return dict(ms=ms.value, launches=n.value, flops=fl.value)


# ----------------------------------------------------------------------------------------
# thin tensor-level wrappers
# ----------------------------------------------------------------------------------------
def ln_modulate(x, shift, scale, rows_per_batch: int, mod_stride: int, eps: float = 1e-6, out=None):
    """x [rows, D]; shift/scale are (views into) modulation rows; see qimg_ln_modulate."""
    _bf16c(x)
    rows, D = x.shape
    out = torch.empty_like(x) if out is None else out
    check(load().qimg_ln_modulate(_p(x), _p(shift), _p(scale), _p(out), rows, D, rows_per_batch, mod_stride, eps,
                                  stream_ptr()), "qimg_ln_modulate")
    return out


def ln_modulate_indexed(x, shift, scale, index, batch: int, rows_per_batch: int, mod_stride: int, eps: float = 1e-6):
    """x [rows, D]; shift / scale: views into a [2 * batch, ...] modulation buffer; index int32 [rows] (0 / 1 per token)."""
    _bf16c(x)
    assert index.dtype == torch.int32 and index.is_cuda and index.is_contiguous() and index.numel() == x.shape[0]
    rows, D = x.shape
    out = torch.empty_like(x)
    check(load().qimg_ln_modulate_indexed(_p(x), _p(shift), _p(scale), _p(out), rows, D, rows_per_batch, mod_stride, eps, _p(index),
                                          batch, stream_ptr()), "qimg_ln_modulate_indexed")
    return out


def select_rows(src, index, batch: int, rows_per_batch: int):
    """src: view [2 * batch, D] (row stride arbitrary, bf16); index int32 [rows] -> [rows, D] with the token's row of `src`."""
    assert src.dtype == torch.bfloat16 and src.stride(1) == 1 and index.dtype == torch.int32
    rows, D = index.numel(), src.shape[1]
    out = torch.empty((rows, D), dtype=torch.bfloat16, device=src.device)
    check(load().qimg_select_rows(_p(src), src.stride(0), _p(index), _p(out), rows, D, rows_per_batch, batch, stream_ptr()),
          "qimg_select_rows")
    return out


def gate_residual(x, y, gate, rows_per_batch: int, gate_stride: int):
    _bf16c(x), _bf16c(y)
    rows, D = x.shape
    check(load().qimg_gate_residual(_p(x), _p(y), _p(gate), rows, D, rows_per_batch, gate_stride, stream_ptr()),
          "qimg_gate_residual")
    return x


def rms_norm(x, w, eps: float = 1e-6):
    _bf16c(x), _bf16c(w)
    rows, D = x.shape
    out = torch.empty_like(x)
    check(load().qimg_rms_norm(_p(x), _p(w), _p(out), rows, D, eps, stream_ptr()), "qimg_rms_norm")
    return out


def linear_small_m(x, W, bias, act_silu: bool, out=None):
    _bf16c(x), _bf16c(W)
    M, K = x.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device) if out is None else out
    check(load().qimg_linear_small_m(_p(x), _p(W), _p(bias), _p(out), M, N, K, out.stride(0), int(act_silu), stream_ptr()),
          "qimg_linear_small_m")
    return out


def timestep_sinusoid(t):
    _bf16c(t)
    out = torch.empty((t.numel(), 256), dtype=torch.bfloat16, device=t.device)
    check(load().qimg_timestep_sinusoid(_p(t), _p(out), t.numel(), stream_ptr()), "qimg_timestep_sinusoid")
    return out


def cfg_euler_step(pos, neg, latents, cfg_scale: float, sigma: float, sigma_next: float):
    """In-place on `latents` [.., 64]."""
    _bf16c(pos), _bf16c(latents)
    if neg is not None:
        _bf16c(neg)
    rows = latents.numel() // latents.shape[-1]
    check(load().qimg_cfg_euler_step(_p(pos), _p(neg), _p(latents), rows, latents.shape[-1], cfg_scale, sigma, sigma_next,
                                     stream_ptr()), "qimg_cfg_euler_step")
    return latents


def cfg_euler_step_dev(pos, neg, latents, cfg_scale: float, sigma_pair: torch.Tensor):
    """`cfg_euler_step` with (sigma_i, sigma_{i+1}) in a device fp32 tensor (CUDA-graph replayable)."""
    _bf16c(pos), _bf16c(latents)
    assert sigma_pair.dtype == torch.float32 and sigma_pair.is_cuda and sigma_pair.numel() >= 2
    rows = latents.numel() // latents.shape[-1]
    check(load().qimg_cfg_euler_step_dev(_p(pos), _p(neg), _p(latents), rows, latents.shape[-1], cfg_scale, _p(sigma_pair),
                                         stream_ptr()), "qimg_cfg_euler_step_dev")
    return latents


def set_nvtx(on: bool):
    """NVTX ranges around every launch family of the engine (Nsight Systems timelines); also env QIMG_NVTX=1."""
    load().qimg_set_nvtx(int(on))


def gemm(problems: list[GemmProblem], epilogue: int):
    arr = (GemmProblem * len(problems))(*problems)
    check(load().qimg_gemm(arr, len(problems), epilogue, stream_ptr()), "qimg_gemm")


def linear(x, W, bias, epilogue: int = EPI_BIAS, out=None):
    """Single-problem convenience: out[M,N] = x[M,K] @ W[N,K]^T + bias (optionally GELU)."""
    _bf16c(x), _bf16c(W), _bf16c(bias)
    M, K = x.shape
    N = W.shape[0]
    out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device) if out is None else out
    p = GemmProblem(A=_p(x), W=_p(W), bias=_p(bias), M=M, N=N, K=K, rows_per_batch=max(M, 1), out=_p(out), ldo=out.stride(0))
    gemm([p], epilogue)
    return out


def fmha_joint(q, k, v, T: int, softmax_scale: float, out_txt=None, out_img=None, mode: int = -1):
    """q,k,v [B,H,S,128] head-major -> (out_txt [B*T, H*128], out_img [B*(S-T), H*128]).
    mode: FMHA_EXACT / FMHA_FAST (fast = guarded by `fmha_overflow`), -1 = process default."""
    for t in (q, k, v):
        _bf16c(t)
    B, H, S, hd = q.shape
    assert hd == 128
    if out_txt is None:
        out_txt = torch.empty((B * T, H * 128), dtype=torch.bfloat16, device=q.device)
    if out_img is None:
        out_img = torch.empty((B * (S - T), H * 128), dtype=torch.bfloat16, device=q.device)
    check(load().qimg_fmha_joint_mode(_p(q), _p(k), _p(v), _p(out_txt), _p(out_img), B, H, S, T, softmax_scale, int(mode),
                                      stream_ptr()), "qimg_fmha_joint")
    return out_txt, out_img


def umma_probe(A, Bm, mode: int):
    D = torch.empty((128, 128), dtype=torch.float32, device=A.device)
    check(load().qimg_umma_probe(_p(A), _p(Bm), _p(D), 128, 128, mode, stream_ptr()), "qimg_umma_probe")
    return D


# ---- peer-memory tensor parallelism (qimg_tp_p2p.cu) ------------------------------------------------------------
def p2p_alloc(nbytes: int) -> int:
    """cudaMalloc'ed zero-filled device buffer that can be exported over CUDA IPC; returns the device pointer."""
    out = C.c_void_p()
    check(load().qimg_p2p_alloc(int(nbytes), C.byref(out)), "qimg_p2p_alloc")
    return int(out.value)


def p2p_free(ptr: int):
    check(load().qimg_p2p_free(C.c_void_p(ptr)), "qimg_p2p_free")


def ipc_get_handle(ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    check(load().qimg_ipc_get_handle(C.c_void_p(ptr), buf), "qimg_ipc_get_handle")
    return bytes(buf.raw)


def ipc_open_handle(handle: bytes) -> int:
    assert len(handle) == 64
    out = C.c_void_p()
    check(load().qimg_ipc_open_handle(C.create_string_buffer(handle, 64), C.byref(out)), "qimg_ipc_open_handle")
    return int(out.value)


def ipc_close_handle(ptr: int):
    check(load().qimg_ipc_close_handle(C.c_void_p(ptr)), "qimg_ipc_close_handle")


# ---- step-cache helpers (TeaCache) ---------------------------------------------------------------------------------
STAGE_PRE, STAGE_BLOCKS, STAGE_POST, STAGE_ALL = 1, 2, 4, 7


def rel_l1_sums(a: torch.Tensor, b: torch.Tensor, sums2: torch.Tensor):
    """sums2 (fp32 [2], device) <- [sum |bf16(a - b)|, sum |b|]."""
    _bf16c(a), _bf16c(b)
    assert a.numel() == b.numel() and sums2.dtype == torch.float32 and sums2.numel() >= 2
    check(load().qimg_rel_l1_sums(_p(a), _p(b), a.numel(), _p(sums2), stream_ptr()), "qimg_rel_l1_sums")


def tea_decide(sums2, n: int, coefficients, thresh: float, accum: torch.Tensor, flag: torch.Tensor, hist, hist_idx: int, force: int):
    """Device-side TeaCache decision (qimg_tea_decide): accum fp64 [1], flag int32 [1], hist fp32 [2 * steps] or None."""
    assert accum.dtype == torch.float64 and flag.dtype == torch.int32
    coef = (C.c_double * 5)(*[float(c) for c in coefficients])
    check(load().qimg_tea_decide(_p(sums2), int(n), coef, float(thresh), _p(accum), _p(flag), _p(hist), int(hist_idx), int(force),
                                 stream_ptr()), "qimg_tea_decide")


def tea_residual(x: torch.Tensor, ori: torch.Tensor, resid: torch.Tensor, flag: torch.Tensor):
    _bf16c(x), _bf16c(ori), _bf16c(resid)
    check(load().qimg_tea_residual(_p(x), _p(ori), _p(resid), x.numel(), _p(flag), stream_ptr()), "qimg_tea_residual")


def bf16_sub(out: torch.Tensor, a: torch.Tensor, b: torch.Tensor):
    _bf16c(out), _bf16c(a), _bf16c(b)
    check(load().qimg_bf16_sub(_p(out), _p(a), _p(b), a.numel(), stream_ptr()), "qimg_bf16_sub")


def bf16_add_inplace(x: torch.Tensor, r: torch.Tensor):
    _bf16c(x), _bf16c(r)
    check(load().qimg_bf16_add_inplace(_p(x), _p(r), x.numel(), stream_ptr()), "qimg_bf16_add_inplace")


# ---- VAE decode (fp32 NHWC; include/qimg_b200.h "VAE decode") ---------------------------------------------------------
def _f32(t: torch.Tensor) -> torch.Tensor:
    assert t.dtype == torch.float32 and t.is_cuda, f"expected a CUDA fp32 tensor, got {t.dtype} on {t.device}"
    return t


def conv2d_nhwc_tf32(x: torch.Tensor, w: torch.Tensor, bias, taps: int, cout: int, res=None, out=None, cin: int | None = None):
    """x [N, H, W, ldx] fp32 (a channel-sliced view with stride-1 channels is fine: ldx = pixel stride, `cin` = channels used);
    w [cout, ldw] packed (co, tap * cin + ci); returns out [N, H, W, cout] = conv + bias (+ res)."""
    _f32(x), _f32(w)
    N, H, W = x.shape[0], x.shape[1], x.shape[2]
    cin = x.shape[3] if cin is None else cin
    ldx = x.stride(2)
    assert x.stride(3) == 1 and x.stride(1) == W * ldx and (N == 1 or x.stride(0) == H * W * ldx), "x must be NHWC with dense pixels"
    if out is None:
        out = torch.empty((N, H, W, cout), dtype=torch.float32, device=x.device)
    assert out.stride(3) == 1 and out.stride(1) == W * out.stride(2)
    ldr = 0
    if res is not None:
        _f32(res)
        assert res.shape[:3] == out.shape[:3] and res.stride(3) == 1 and res.stride(1) == W * res.stride(2)
        ldr = res.stride(2)
    check(load().qimg_conv2d_nhwc_tf32(_p(x), ldx, _p(w), w.stride(0), _p(bias), _p(res), ldr, _p(out), out.stride(2), N, H, W,
                                       cin, cout, taps, stream_ptr()), "qimg_conv2d_nhwc_tf32")
    return out


def set_vae_conv_variant(variant: int):
    check(load().qimg_set_vae_conv_variant(int(variant)), "qimg_set_vae_conv_variant")


def conv2d_down2_nhwc_tf32(x: torch.Tensor, w: torch.Tensor, bias, cout: int):
    """ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2): x [N, H, W, C] contiguous fp32 -> [N, H/2, W/2, cout]."""
    _f32(x), _f32(w)
    assert x.is_contiguous()
    N, H, W, C_ = x.shape
    out = torch.empty((N, H // 2, W // 2, cout), dtype=torch.float32, device=x.device)
    check(load().qimg_conv2d_down2_nhwc_tf32(_p(x), C_, _p(w), w.stride(0), _p(bias), _p(out), cout, N, H, W, C_, cout, stream_ptr()),
          "qimg_conv2d_down2_nhwc_tf32")
    return out


def vae_image_to_nhwc(img: torch.Tensor):
    """img [N, C <= 32, H, W] fp32 NCHW -> [N, H, W, 32] NHWC (zero-padded channels)."""
    _f32(img)
    assert img.is_contiguous()
    N, C_, H, W = img.shape
    out = torch.empty((N, H, W, 32), dtype=torch.float32, device=img.device)
    check(load().qimg_vae_image_to_nhwc(_p(img), _p(out), N, C_, H, W, stream_ptr()), "qimg_vae_image_to_nhwc")
    return out


def vae_rms_act(x: torch.Tensor, gamma: torch.Tensor, silu: bool, out=None):
    _f32(x), _f32(gamma)
    assert x.is_contiguous()
    C_ = x.shape[-1]
    out = torch.empty_like(x) if out is None else out
    check(load().qimg_vae_rms_act(_p(x), _p(gamma), _p(out), x.numel() // C_, C_, 1 if silu else 0, stream_ptr()), "qimg_vae_rms_act")
    return out


def vae_upsample2x(x: torch.Tensor):
    _f32(x)
    assert x.is_contiguous()
    N, H, W, C_ = x.shape
    out = torch.empty((N, 2 * H, 2 * W, C_), dtype=torch.float32, device=x.device)
    check(load().qimg_vae_upsample2x(_p(x), _p(out), N, H, W, C_, stream_ptr()), "qimg_vae_upsample2x")
    return out


def vae_post_quant(z: torch.Tensor, w: torch.Tensor, b: torch.Tensor):
    """z [N, 16, H, W] fp32 NCHW -> [N, H, W, 32] NHWC (channels 16.. zero)."""
    _f32(z), _f32(w), _f32(b)
    assert z.is_contiguous() and w.is_contiguous()
    N, Z, H, W = z.shape
    out = torch.empty((N, H, W, 32), dtype=torch.float32, device=z.device)
    check(load().qimg_vae_post_quant(_p(z), _p(w), _p(b), _p(out), N, H, W, Z, stream_ptr()), "qimg_vae_post_quant")
    return out


def vae_conv_out(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, uint8: bool = False):
    """x [N, H, W, 96] (normalised + SiLU) -> clamp(conv3x3 -> 3 channels); w packed [3, 9, 96].  Returns NCHW fp32
    [N, 3, H, W], or with `uint8` the post-processed NHWC uint8 image [N, H, W, 3]."""
    _f32(x), _f32(w), _f32(b)
    assert x.is_contiguous() and w.is_contiguous()
    N, H, W, C_ = x.shape
    if uint8:
        out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device)
        check(load().qimg_vae_conv_out(_p(x), _p(w), _p(b), None, _p(out), N, H, W, C_, stream_ptr()), "qimg_vae_conv_out")
    else:
        out = torch.empty((N, 3, H, W), dtype=torch.float32, device=x.device)
        check(load().qimg_vae_conv_out(_p(x), _p(w), _p(b), _p(out), None, N, H, W, C_, stream_ptr()), "qimg_vae_conv_out")
    return out


def vae_softmax_rows(s: torch.Tensor, scale: float):
    _f32(s)
    assert s.dim() == 2 and s.stride(1) == 1
    check(load().qimg_vae_softmax_rows(_p(s), s.shape[0], s.shape[1], s.stride(0), float(scale), stream_ptr()), "qimg_vae_softmax_rows")
    return s


def vae_transpose(x: torch.Tensor):
    """x [rows, cols] fp32 (row stride arbitrary) -> contiguous [cols, rows]."""
    _f32(x)
    assert x.dim() == 2 and x.stride(1) == 1
    out = torch.empty((x.shape[1], x.shape[0]), dtype=torch.float32, device=x.device)
    check(load().qimg_vae_transpose(_p(x), x.stride(0), _p(out), x.shape[0], x.shape[1], stream_ptr()), "qimg_vae_transpose")
    return out
