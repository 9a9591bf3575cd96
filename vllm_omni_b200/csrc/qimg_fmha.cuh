// Joint (text + image) non-causal attention for the Qwen-Image MMDiT block, sm_100a: shared definitions.
//
// Replaces Attention.forward -> SDPAImpl.forward (reference attention/layer.py:54-70,
// backends/sdpa.py:46-66) plus the torch.cat / permute / split copies around it
// (qwen_image_transformer.py:414-416, 444-449): Q/K/V arrive head-major
// [B, H, S = T + S_img, 128] (written directly by the QKV GEMM epilogue) and the output is
// written straight into the two per-stream [rows, H*128] buffers the out-projections read.
//
// Two pipelines ship (round 1 explored seven; the losers were deleted, see profiles/r01_fmha_trace.md):
//   qimg_fmha4.cuh  fmha_joint_kernel_v7  EXACT: every KV tile's row maximum is reduced before its exponentials
//   qimg_fmha6.cuh  fmha_joint_kernel_v9  FAST (default): exponentials against the row's running reference maximum,
//                                         guarded by a device-side overflow flag (qimg_fmha_overflow)
// Common structure, one CTA per (256 query rows, batch*head), 576 threads:
//   warp 0     : TMA producer (Q once; K and V tiles through separate 2-stage mbarrier rings)
//   warp 1     : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-17 : softmax, TWO threads per query row (8 warps per 128-row query tile)
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each.  P (bf16) overwrites part of its S tile and
// feeds the P*V MMA as the TMEM A operand; V is consumed in its natural [kv, hd] layout as an MN-major B
// operand.  The MMA issue order QK0(j) PV1(j-1) QK1(j) PV0(j) keeps the tensor pipe busy while the other
// tile's softmax runs.  Online softmax with lazy (warp-uniform, threshold 2^8) rescaling of the O accumulator.
#pragma once

#include <type_traits>

#include "qimg_common.cuh"

namespace qimg {

// Cycle tracing (qimg_set_fmha_trace, tools/fmha_trace.py) is compiled in only with -DQIMG_FMHA_TRACE
// (QIMG_FMHA_TRACE=1 python -m vllm_omni_b200.build): the counters cost registers in the softmax warps.
#ifdef QIMG_FMHA_TRACE
constexpr bool kFmhaTrace = true;
#else
constexpr bool kFmhaTrace = false;
#endif

constexpr int FMHA_THREADS = 320;
constexpr int FMHA_KS = 2;
constexpr int FMHA_VS = 2;
constexpr int FMHA_TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column SW128 slabs of 16 KB
constexpr int FMHA_SMEM_BYTES = (2 + FMHA_KS + FMHA_VS) * FMHA_TILE_BYTES + 1024 + 256;

struct FmhaParams {
  bf16* out_txt;  // [B*T, H*128]
  bf16* out_img;  // [B*S_img, H*128]
  int B, H, S, T;
  float scale_log2;  // softmax_scale * log2(e)
  long long* trace;  // optional (qimg_set_fmha_trace): per-phase cycle counters of CTA 200
  // sequence parallelism (Ulysses): this rank computed H LOCAL heads over ALL rows; output rows go to the rank that owns
  // them — the all-to-all behind the attention (reference attention/parallel/ulysses.py:135) as peer stores of the epilogue.
  // sp_out_img/txt[o]: owner o's [own rows, H * sp_size * 128] buffers; rows are split balanced over sp_size owners.
  int sp_size, sp_rank, sp_rows_img, sp_rows_txt;
  bf16* sp_out_img[8];
  bf16* sp_out_txt[8];
  int single_tile;   // 1: one 128-row query tile per CTA (grid = tiles x B x H) — chosen by the launcher when that still
                     // fits one wave, e.g. 3 local heads under TP=8: 99 half-size CTAs instead of 51 full-size ones
  const int* skip;   // optional device predicate: non-zero -> exit at once (step-cache reuse)
  int* overflow;     // fast pipeline: set to 1 when a score exceeded the row's reference maximum by > 2^FMHA_OVF_LOG2
};

// The fast pipeline exponentiates tile j >= 1 against a reference that may lag the true row maximum; exp2 arguments
// up to 100 stay finite in fp32 / bf16 and lose nothing (floating point), beyond that the launch is flagged and the
// caller recomputes with the exact pipeline (pipeline_qwen_image.py::_denoise, qimg_fmha_overflow).
constexpr float FMHA_OVF_LOG2 = 100.0f;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32x2 helpers (Blackwell FFMA2 / FADD2 / FMNMX3): halve the issue slots of the softmax ----
__device__ __forceinline__ uint64_t pack_f32x2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(uint64_t v, uint32_t& lo, uint32_t& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ float max3_f32(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ uint64_t splat_f32x2(float v) { return pack_f32x2(__float_as_uint(v), __float_as_uint(v)); }

// 2^x for a pair on the FMA pipe (no MUFU): round-to-nearest split x = n + f, f in [-0.5, 0.5], degree-3
// minimax polynomial for 2^f (max rel. error 7.5e-5, far below the bf16 rounding of P), exponent add.
// Softmax exponentials are XU (MUFU.EX2, 16/clk/SM) bound at head_dim 128; FMHA_POLY_MASK routes a
// fraction of the pairs here so XU and FMA pipes share the load (FlashAttention-4's trick).
__device__ __forceinline__ uint64_t exp2_poly_f32x2(uint64_t x) {
  uint32_t xl, xh;
  unpack_f32x2(x, xl, xh);
  xl = __float_as_uint(fmaxf(__uint_as_float(xl), -125.0f));
  xh = __float_as_uint(fmaxf(__uint_as_float(xh), -125.0f));
  x = pack_f32x2(xl, xh);
  const uint64_t magic = splat_f32x2(12582912.0f);  // 1.5 * 2^23: low mantissa bits of t hold round(x)
  const uint64_t t = add_f32x2(x, magic);
  const uint64_t n = add_f32x2(t, splat_f32x2(-12582912.0f));
  const uint64_t f = fma_f32x2(n, splat_f32x2(-1.0f), x);
  uint64_t p = fma_f32x2(f, splat_f32x2(0.05517164617776871f), splat_f32x2(0.2426111251115799f));
  p = fma_f32x2(p, f, splat_f32x2(0.6932609677314758f));
  p = fma_f32x2(p, f, splat_f32x2(0.9999280571937561f));
  uint32_t pl, ph, tl, th;
  unpack_f32x2(p, pl, ph);
  unpack_f32x2(t, tl, th);
  return pack_f32x2(pl + (tl << 23), ph + (th << 23));
}

// Destination of the 16-byte chunk c16 of output row (batch b, joint position pos, local head h).
__device__ __forceinline__ bf16* fmha_out_ptr(const FmhaParams& prm, int b, int h, int pos, int c16) {
  const int S_img = prm.S - prm.T;
  const bool txt = pos < prm.T;
  if (prm.sp_size <= 1) {
    const int D = prm.H * 128;
    bf16* row = txt ? prm.out_txt + ((size_t)b * prm.T + pos) * D : prm.out_img + ((size_t)b * S_img + (pos - prm.T)) * D;
    return row + h * 128 + c16 * 8;
  }
  const int m = txt ? b * prm.T + pos : b * S_img + (pos - prm.T);  // global row of its stream
  const int M = txt ? prm.sp_rows_txt : prm.sp_rows_img;
  const int base = M / prm.sp_size, extra = M % prm.sp_size, cut = extra * (base + 1);
  int o, li;
  if (m < cut) {
    o = m / (base + 1);
    li = m - o * (base + 1);
  } else {
    o = extra + (m - cut) / base;
    li = (m - cut) - (o - extra) * base;
  }
  const int Dtot = prm.H * prm.sp_size * 128;
  bf16* buf = txt ? prm.sp_out_txt[o] : prm.sp_out_img[o];
  return buf + (size_t)li * Dtot + (prm.sp_rank * prm.H + h) * 128 + c16 * 8;
}

// CTA -> (batch*head, query-tile pair).  S = 4224 gives 16 full pairs + 1 "half" pair (second tile beyond S) per head.
// Full pairs run head-major so the 16 CTAs of a head stream the same K/V through L2 together.  The half pair of head h
// is slotted FMHA_HALF_LAG heads later (its K/V is still L2-resident: round 1 ran all half pairs at the very end and
// re-read every head's K/V from DRAM, 514 MB instead of 311 MB per launch, profiles/r01_ncu_fmha_v9.csv); the last
// FMHA_HALF_LAG heads' half pairs still run last, so the grid ends on short CTAs.
constexpr int FMHA_HALF_LAG = 8;
struct FmhaWork {
  int bh, pair_idx;
};
__device__ __forceinline__ FmhaWork fmha_decode_cta(int i, int S, int n_bh) {
  const int pairs = (S + 255) / 256;
  const int F = S / 256 + ((S % 256) > 128 ? 1 : 0);  // pairs with both tiles (at least partly) in range
  FmhaWork w;
  if (F == pairs) {  // no half pair
    w.bh = i / pairs;
    w.pair_idx = i - w.bh * pairs;
    return w;
  }
  const int lag = n_bh < FMHA_HALF_LAG ? n_bh : FMHA_HALF_LAG;
  if (i < lag * F) {
    w.bh = i / F;
    w.pair_idx = i - w.bh * F;
    return w;
  }
  const int i2 = i - lag * F;
  const int g = i2 / (F + 1), s = i2 - g * (F + 1);
  if (lag + g < n_bh) {
    if (s < F) {
      w.bh = lag + g;
      w.pair_idx = s;
    } else {
      w.bh = g;  // half pair of the head that ran `lag` heads ago
      w.pair_idx = F;
    }
    return w;
  }
  w.bh = n_bh - lag + (i2 - (n_bh - lag) * (F + 1));
  w.pair_idx = F;
  return w;
}

}  // namespace qimg
