// Joint (text + image) non-causal attention for the Qwen-Image MMDiT block, sm_100a.
//
// Replaces Attention.forward -> SDPAImpl.forward (reference attention/layer.py:54-70,
// backends/sdpa.py:46-66) plus the torch.cat / permute / split copies around it
// (qwen_image_transformer.py:414-416, 444-449): Q/K/V arrive head-major
// [B, H, S = T + S_img, 128] (written directly by the QKV GEMM epilogue) and the output is
// written straight into the two per-stream [rows, H*128] buffers the out-projections read.
//
// FA4-style structure, one CTA per (256 query rows, batch*head), 320 threads:
//   warp 0    : TMA producer (Q once; K and V tiles through separate 2-stage mbarrier rings)
//   warp 1    : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5 : softmax warpgroup for query tile 0 (one thread per query row)
//   warps 6-9 : softmax warpgroup for query tile 1
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each.  P (bf16) overwrites the
// first 64 columns of its S tile and feeds the P*V MMA as the TMEM A operand; V is consumed
// in its natural [kv, hd] layout as an MN-major B operand.  The MMA issue order
// QK0(j) PV1(j-1) QK1(j) PV0(j) keeps the tensor pipe busy while the other tile's softmax runs.
// Online softmax with lazy (warp-uniform, threshold 2^8) rescaling of the O accumulator.
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
  long long* trace;  // optional (qimg_set_fmha_trace): per-phase cycle counters of CTA 200 (pipelines 0 and 4)
};

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

// Measured on B200 (profiles/r01_fmha_v3): with 37.5 % of the pairs on the polynomial the XU pipe sat at 31 %
// and the kernel was bound by softmax issue slots / latency, not by MUFU, so the default routes everything
// to MUFU.EX2 (fewest instructions per element); the polynomial stays available for head sizes / chips where
// the XU pipe saturates first.
#ifndef FMHA_POLY_MASK
#define FMHA_POLY_MASK 0x00u  // bit k set => pair (k mod 8) of every 8 pairs uses the FMA-pipe polynomial
#endif

template <uint32_t POLY_MASK, bool PINGPONG>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_joint_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FmhaParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 tiles
  uint8_t* sK = smem + 2 * FMHA_TILE_BYTES;             // KS tiles
  uint8_t* sV = sK + FMHA_KS * FMHA_TILE_BYTES;         // VS tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FMHA_VS * FMHA_TILE_BYTES);
  uint64_t* q_full = bars;              // [1]
  uint64_t* k_full = bars + 1;          // [KS]
  uint64_t* k_empty = k_full + FMHA_KS;
  uint64_t* v_full = k_empty + FMHA_KS;
  uint64_t* v_empty = v_full + FMHA_VS;
  uint64_t* s_full = v_empty + FMHA_VS;  // [2]
  uint64_t* p_ready = s_full + 2;        // [2]
  uint64_t* o_full = p_ready + 2;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // 1-D grid, remapped so that the query-tile pairs whose SECOND tile lies completely beyond S run last and
  // skip that tile (S = 4224 -> 16 full pairs + 1 half pair per head; 1536 full + 96 half CTAs fill 148 SMs in
  // ~11.1 instead of 12 CTA-times)
  const int full_pairs = prm.S / 256 + ((prm.S % 256) > 128 ? 1 : 0);
  const int n_bh = prm.B * prm.H;
  int bh, pair_idx;
  if ((int)blockIdx.x < full_pairs * n_bh) {
    bh = blockIdx.x / full_pairs;
    pair_idx = blockIdx.x - bh * full_pairs;
  } else {
    bh = blockIdx.x - full_pairs * n_bh;
    pair_idx = full_pairs;
  }
  const int q_row0 = pair_idx * 256;
  const bool two = q_row0 + 128 < prm.S;  // is the second query tile (partly) in range?
  const int n_kv = (prm.S + 127) / 128;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FMHA_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, (two ? 2 : 1) * FMHA_TILE_BYTES);
      for (int t = 0; t < (two ? 2 : 1); ++t)
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sQ + t * FMHA_TILE_BYTES + s * 16384, &tmQ, q_full, s * 64, q_row0 + t * 128, bh);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int ks = j % FMHA_KS, vs = j % FMHA_VS;
      mbar_wait(&k_empty[ks], ((j / FMHA_KS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[ks], FMHA_TILE_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sK + ks * FMHA_TILE_BYTES + s * 16384, &tmK, &k_full[ks], s * 64, j * 128, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[vs], ((j / FMHA_VS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[vs], FMHA_TILE_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sV + vs * FMHA_TILE_BYTES + s * 16384, &tmV, &v_full[vs], s * 64, j * 128, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);  // B (=V) is MN-major
    const uint32_t tS[2] = {tmem_base + 0, tmem_base + 128};
    const uint32_t tO[2] = {tmem_base + 256, tmem_base + 384};
    auto issue_qk = [&](int t, int ks) {
      const uint32_t qa = smem_u32(sQ + t * FMHA_TILE_BYTES);
      const uint32_t ka = smem_u32(sK + ks * FMHA_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tS[t], make_kmajor_sw128_desc(qa + off), make_kmajor_sw128_desc(ka + off), IDESC_QK, k != 0);
      }
    };
    auto issue_pv = [&](int t, int vs, bool accumulate) {
      const uint32_t va = smem_u32(sV + vs * FMHA_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        // A = P (bf16 pairs, 8 TMEM columns per K=16 step); B = V rows [16k, 16k+16) x 128 (MN-major)
        umma_ts(tO[t], tS[t] + k * 8, make_mnmajor_sw128_desc(va + k * 2048, 16384), IDESC_PV,
                (accumulate || k != 0) ? 1u : 0u);
      }
    };
    const bool tr = kFmhaTrace && prm.trace != nullptr && blockIdx.x == 200;
    long long w_k = 0, w_p1 = 0, w_v = 0, w_p0 = 0, tt = 0;
    mbar_wait(q_full, 0);
    const long long t_begin = kFmhaTrace ? clock64() : 0;
    for (int j = 0; j < n_kv; ++j) {
      const int ks = j % FMHA_KS;
      if (tr) tt = clock64();
      mbar_wait(&k_full[ks], (j / FMHA_KS) & 1);
      if (tr) w_k += clock64() - tt;
      tc_fence_after();
      if (elect_one()) {
        issue_qk(0, ks);
        umma_commit(&s_full[0]);
      }
      __syncwarp();
      if (two && j > 0) {
        if (tr) tt = clock64();
        mbar_wait(&p_ready[1], (j - 1) & 1);
        if (tr) w_p1 += clock64() - tt;
        tc_fence_after();
        if (elect_one()) issue_pv(1, (j - 1) % FMHA_VS, j - 1 > 0);
        __syncwarp();
      }
      if (elect_one()) {
        if (j > 0) umma_commit(&v_empty[(j - 1) % FMHA_VS]);  // V(j-1): PV0(j-1) and PV1(j-1) are both issued
        if (two) {
          issue_qk(1, ks);
          umma_commit(&s_full[1]);
        }
        umma_commit(&k_empty[ks]);
      }
      __syncwarp();
      const int vs = j % FMHA_VS;
      if (tr) tt = clock64();
      mbar_wait(&v_full[vs], (j / FMHA_VS) & 1);
      if (tr) w_v += clock64() - tt, tt = clock64();
      mbar_wait(&p_ready[0], j & 1);
      if (tr) w_p0 += clock64() - tt;
      tc_fence_after();
      if (elect_one()) issue_pv(0, vs, j > 0);
      __syncwarp();
    }
    if (two) {
      mbar_wait(&p_ready[1], (n_kv - 1) & 1);
      tc_fence_after();
    }
    if (elect_one()) {
      if (two) issue_pv(1, (n_kv - 1) % FMHA_VS, n_kv - 1 > 0);
      umma_commit(&v_empty[(n_kv - 1) % FMHA_VS]);
      umma_commit(&o_full[0]);
      umma_commit(&o_full[1]);
    }
    __syncwarp();
    if (tr && lane == 0) {
      prm.trace[0] = clock64() - t_begin;
      prm.trace[1] = w_k;
      prm.trace[2] = w_p1;
      prm.trace[3] = w_v;
      prm.trace[4] = w_p0;
      prm.trace[5] = n_kv;
    }
  } else {
    // ===================== softmax / correction / output warps =====================
    const int t = (warp - 2) >> 2;  // query tile handled by this warpgroup
    if (t == 0 || two) {
    const bool pingpong = PINGPONG && two;
    const int q = warp & 3;         // TMEM lane quarter
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * 128;
    const float c = prm.scale_log2;
    float m_used = -INFINITY;  // row max (raw score units) the exponentials are referenced to
    float l = 0.f;             // running row sum
    // Exponential phases of the two query tiles strictly alternate (named barriers 1 / 2, 256 threads): the
    // two warpgroups would otherwise run their MUFU-heavy phases at the same time, halving each other's
    // XU throughput and then both waiting for the tensor pipe.  Alternation keeps tile 0's softmax under
    // tile 1's MMAs and vice versa (the issue order QK0 PV1 QK1 PV0 assumes exactly that).
    if (pingpong && t == 1) named_bar_arrive(1, 256);
    const bool tr = kFmhaTrace && prm.trace != nullptr && blockIdx.x == 200 && q == 0;
    long long w_s = 0, w_ld = 0, w_mx = 0, w_pp = 0, w_ex = 0, w_tl = 0, tt = 0;
    const long long t_begin = kFmhaTrace ? clock64() : 0;
    for (int j = 0; j < n_kv; ++j) {
      if (tr) tt = clock64();
      mbar_wait(&s_full[t], j & 1);
      if (tr) w_s += clock64() - tt, tt = clock64();
      tc_fence_after();
      const int kv_valid = prm.S - j * 128;  // < 128 only on a ragged last tile
      // The body is instantiated twice; only the ragged last KV tile pays for the 128 compare/selects
      // of the -inf masking (as one predicated block they would be executed on every tile).
      auto softmax_tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // ---- one TMEM round trip: the whole 128-wide score row lives in registers ----
        uint32_t r[128];
  #pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(tS + cc * 32, r + cc * 32);
        tmem_ld_wait();
        if (tr) w_ld += clock64() - tt, tt = clock64();
        if (MASKED) {
  #pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) r[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
  #pragma unroll
        for (int i = 0; i < 128; i += 8) {
          mx0 = max3_f32(mx0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
          mx1 = max3_f32(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
          mx2 = max3_f32(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
          mx3 = max3_f32(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        if (j == 0) {
          m_used = mx;
        } else {
          const float m_new = fmaxf(m_used, mx);
          const bool need = (m_new - m_used) * c > 8.0f;
          if (__any_sync(0xffffffffu, need)) {
            // rescale O and l to the new reference max (PV(j-1) of this tile is complete: s_full(j)
            // was committed after it in the in-order tensor pipe)
            const float f = ex2_approx((m_used - m_new) * c);
            l *= f;
  #pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {
              uint32_t o[32];
              tmem_ld_32x32b_x32(tO + cc * 32, o);
              tmem_ld_wait();
  #pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_32x32b_x32(tO + cc * 32, o);
            }
            tmem_st_wait();
            m_used = m_new;
          }
        }
        // ---- P = exp2(s*c - m*c) (masked columns give exp2(-inf) = 0), row sum, bf16 P -> TMEM ----
        if (tr) w_mx += clock64() - tt, tt = clock64();
        if (pingpong) named_bar_sync(1 + t, 256);  // my turn on the XU pipe
        if (tr) w_pp += clock64() - tt, tt = clock64();
        const uint64_t c2 = splat_f32x2(c), nmc2 = splat_f32x2(-m_used * c);
        uint64_t la = 0, lb = 0;  // two packed partial row sums (bit pattern 0 = +0.0f pairs)
  #pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t pk[16];
  #pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = cc * 16 + i;  // pair index 0..63
            const uint64_t x = fma_f32x2(pack_f32x2(r[2 * k], r[2 * k + 1]), c2, nmc2);
            uint64_t p;
            if ((POLY_MASK >> (k & 7)) & 1u) {
              p = exp2_poly_f32x2(x);
            } else {
              uint32_t xl, xh;
              unpack_f32x2(x, xl, xh);
              p = pack_f32x2(__float_as_uint(ex2_approx(__uint_as_float(xl))), __float_as_uint(ex2_approx(__uint_as_float(xh))));
            }
            if (i & 1) lb = add_f32x2(lb, p); else la = add_f32x2(la, p);
            uint32_t pl, ph;
            unpack_f32x2(p, pl, ph);
            pk[i] = pack_bf16x2(__uint_as_float(pl), __uint_as_float(ph));
          }
          tmem_st_32x32b_x16(tS + cc * 16, pk);
        }
        if (tr) w_ex += clock64() - tt, tt = clock64();
        if (pingpong && !(t == 1 && j == n_kv - 1)) named_bar_arrive(1 + (t ^ 1), 256);  // hand the XU pipe over
        {
          uint32_t a0, a1, b0, b1;
          unpack_f32x2(la, a0, a1);
          unpack_f32x2(lb, b0, b1);
          l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
        }
      };
      if (kv_valid < 128) softmax_tile(std::true_type{});
      else softmax_tile(std::false_type{});
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);
      if (tr) w_tl += clock64() - tt;
    }
    if (tr && lane == 0) {
      long long* o = prm.trace + 8 + t * 8;
      o[0] = clock64() - t_begin;
      o[1] = w_s;
      o[2] = w_ld;
      o[3] = w_mx;
      o[4] = w_pp;
      o[5] = w_ex;
      o[6] = w_tl;
    }
    // ---- final: O / l -> bf16 -> smem (this tile's Q buffer is free now) -> coalesced stores ----
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    uint8_t* stg = sQ + t * FMHA_TILE_BYTES;  // 128 rows x 256 B
    const int row = q * 32 + lane;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tO + cc * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(r[g * 8 + 0]) * inv_l, __uint_as_float(r[g * 8 + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(r[g * 8 + 2]) * inv_l, __uint_as_float(r[g * 8 + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(r[g * 8 + 4]) * inv_l, __uint_as_float(r[g * 8 + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(r[g * 8 + 6]) * inv_l, __uint_as_float(r[g * 8 + 7]) * inv_l);
        const int c16 = cc * 4 + g;  // 16-byte chunk index within the 256 B row
        *reinterpret_cast<uint4*>(stg + row * 256 + ((c16 ^ (row & 7)) << 4)) = v;
      }
    }
    __syncwarp();
    const int b = bh / prm.H, h = bh - b * prm.H;
    const int D = prm.H * 128;
    const int S_img = prm.S - prm.T;
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
      const int rr = q * 32 + it * 2 + (lane >> 4);
      const int c16 = lane & 15;
      const int pos = q_row0 + t * 128 + rr;
      if (pos < prm.S) {
        uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 256 + ((c16 ^ (rr & 7)) << 4));
        bf16* dst = (pos < prm.T) ? prm.out_txt + ((size_t)b * prm.T + pos) * D
                                  : prm.out_img + ((size_t)b * S_img + (pos - prm.T)) * D;
        stg_v4(dst + h * 128 + c16 * 8, v);
      }
    }
    }  // active tile
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace qimg
