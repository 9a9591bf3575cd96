// Joint attention, second-generation pipeline (FMHA "v5"): same interface and data layout as
// fmha_joint_kernel (qimg_fmha.cuh) but the KV tile is 64 rows and every query tile owns TWO score
// buffers in TMEM, so Q*K^T of KV tile j+1 (and j+2) is issued BEFORE the softmax of tile j has
// finished.  The first-generation kernel (128-row KV tiles, one S buffer per query tile) was bound by
// the latency of the chain  S(j) -> softmax -> P(j) -> PV(j) -> QK(j+1)  (profiles/r01_fmha_v3:
// tensor pipe 50 % active, softmax warps 36 % of their samples waiting for S, MMA thread waiting for P);
// with double-buffered S only softmax THROUGHPUT has to keep up with the tensor pipe.
//
// TMEM (512 columns): S[t][b] at (2t+b)*64 (t = query tile 0/1, b = j & 1), O[t] at 256 + 128 t.
// P(j) (bf16) overwrites the first 32 columns of S[t][j&1] and is the TMEM A operand of P*V.
// MMA issue order per KV step j:  PV0(j) QK0(j+2) PV1(j) QK1(j+2)   (QK(j+2) reuses buffer j&1, freed by PV(j)).
#pragma once

#include <type_traits>

#include "qimg_fmha.cuh"

namespace qimg {

constexpr int FMHA2_KV = 64;
constexpr int FMHA2_KS = 4;
constexpr int FMHA2_VS = 4;
constexpr int FMHA2_Q_BYTES = 128 * 128 * 2;   // 32 KB per query tile (two 64-col SW128 slabs of 16 KB)
constexpr int FMHA2_KV_BYTES = 64 * 128 * 2;   // 16 KB per K or V tile (two slabs of 8 KB)
constexpr int FMHA2_SMEM_BYTES = 2 * FMHA2_Q_BYTES + (FMHA2_KS + FMHA2_VS) * FMHA2_KV_BYTES + 1024 + 512;

#ifndef FMHA2_POLY_MASK
#define FMHA2_POLY_MASK 0x11u  // pairs {0,4} of every 8 -> 25 % of the exponentials on the FMA pipe
#endif

template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_joint_kernel_v5(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FmhaParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * FMHA2_Q_BYTES;
  uint8_t* sV = sK + FMHA2_KS * FMHA2_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FMHA2_VS * FMHA2_KV_BYTES);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = bars + 1;                  // [KS]
  uint64_t* k_empty = k_full + FMHA2_KS;        // [KS]
  uint64_t* v_full = k_empty + FMHA2_KS;        // [VS]
  uint64_t* v_empty = v_full + FMHA2_VS;        // [VS]
  uint64_t* s_full = v_empty + FMHA2_VS;        // [2 tiles][2 buffers]
  uint64_t* p_ready = s_full + 4;               // [2][2]
  uint64_t* pv_done = p_ready + 4;              // [2]   one phase per PV(t, j)
  uint64_t* o_full = pv_done + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q_row0 = blockIdx.x * 256;
  const int n_kv = (prm.S + FMHA2_KV - 1) / FMHA2_KV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA2_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FMHA2_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&pv_done[i], 1);
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
      mbar_arrive_expect_tx(q_full, 2 * FMHA2_Q_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sQ + t * FMHA2_Q_BYTES + s * 16384, &tmQ, q_full, s * 64, q_row0 + t * 128, bh);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int ks = j % FMHA2_KS, vs = j % FMHA2_VS;
      mbar_wait(&k_empty[ks], ((j / FMHA2_KS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[ks], FMHA2_KV_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sK + ks * FMHA2_KV_BYTES + s * 8192, &tmK, &k_full[ks], s * 64, j * FMHA2_KV, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[vs], ((j / FMHA2_VS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[vs], FMHA2_KV_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sV + vs * FMHA2_KV_BYTES + s * 8192, &tmV, &v_full[vs], s * 64, j * FMHA2_KV, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, FMHA2_KV, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);  // B (=V) is MN-major
    auto issue_qk = [&](int t, int b, int ks) {
      const uint32_t qa = smem_u32(sQ + t * FMHA2_Q_BYTES);
      const uint32_t ka = smem_u32(sK + ks * FMHA2_KV_BYTES);
      const uint32_t d = tmem_base + (2 * t + b) * 64;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        umma_ss(d, make_kmajor_sw128_desc(qa + (k >> 2) * 16384 + (k & 3) * 32),
                make_kmajor_sw128_desc(ka + (k >> 2) * 8192 + (k & 3) * 32), IDESC_QK, k != 0);
      }
    };
    auto issue_pv = [&](int t, int b, int vs, bool accumulate) {
      const uint32_t va = smem_u32(sV + vs * FMHA2_KV_BYTES);
      const uint32_t d = tmem_base + 256 + t * 128;
      const uint32_t p = tmem_base + (2 * t + b) * 64;
#pragma unroll
      for (int k = 0; k < FMHA2_KV / 16; ++k) {
        // A = P (bf16 pairs, 8 TMEM columns per K=16 step); B = V rows [16k,16k+16) x 128 (MN-major, slabs 8 KB apart)
        umma_ts(d, p + k * 8, make_mnmajor_sw128_desc(va + k * 2048, 8192), IDESC_PV, (accumulate || k != 0) ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    // prologue: scores of the first two KV tiles for both query tiles
    for (int jj = 0; jj < 2 && jj < n_kv; ++jj) {
      const int ks = jj % FMHA2_KS;
      mbar_wait(&k_full[ks], (jj / FMHA2_KS) & 1);
      tc_fence_after();
      if (elect_one()) {
        for (int t = 0; t < 2; ++t) {
          issue_qk(t, jj & 1, ks);
          umma_commit(&s_full[2 * t + (jj & 1)]);
        }
        umma_commit(&k_empty[ks]);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1, vs = j % FMHA2_VS;
      const int jn = j + 2, ksn = jn % FMHA2_KS;
      mbar_wait(&v_full[vs], (j / FMHA2_VS) & 1);
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_ready[2 * t + b], (j >> 1) & 1);
        if (jn < n_kv && t == 0) mbar_wait(&k_full[ksn], (jn / FMHA2_KS) & 1);
        tc_fence_after();
        if (elect_one()) {
          issue_pv(t, b, vs, j > 0);
          umma_commit(&pv_done[t]);
          if (jn < n_kv) {
            issue_qk(t, b, ksn);  // reuses S[t][b]: P(j) was consumed by the PV issued just above (in-order pipe)
            umma_commit(&s_full[2 * t + b]);
            if (t == 1) umma_commit(&k_empty[ksn]);
          }
          if (t == 1) umma_commit(&v_empty[vs]);
        }
        __syncwarp();
      }
    }
    if (elect_one()) {
      umma_commit(&o_full[0]);
      umma_commit(&o_full[1]);
    }
    __syncwarp();
  } else {
    // ===================== softmax / correction / output warps =====================
    const int t = (warp - 2) >> 2;  // query tile handled by this warpgroup
    const int q = warp & 3;         // TMEM lane quarter
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tO = tmem_base + lane_off + 256 + t * 128;
    const float c = prm.scale_log2;
    float m_used = -INFINITY;  // row max (raw score units) the exponentials are referenced to
    float l = 0.f;             // running row sum
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      const uint32_t tS = tmem_base + lane_off + (2 * t + b) * 64;
      mbar_wait(&s_full[2 * t + b], (j >> 1) & 1);
      tc_fence_after();
      const int kv_valid = prm.S - j * FMHA2_KV;  // < 64 only on a ragged last tile
      auto softmax_tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        uint32_t r[64];
        tmem_ld_32x32b_x32(tS, r);
        tmem_ld_32x32b_x32(tS + 32, r + 32);
        tmem_ld_wait();
        if (MASKED) {
#pragma unroll
          for (int i = 0; i < 64; ++i)
            if (i >= kv_valid) r[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; i += 8) {
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
            // Lazy rescale of O and l.  QK(j) was issued BEFORE PV(j-1) in this pipeline, so wait for
            // PV(t, j-1) explicitly: pv_done[t] completes one phase per PV, and at this point phases
            // 0..j-2 are certainly complete, so the parity of phase j-1 is unambiguous.
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
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
        const uint64_t c2 = splat_f32x2(c), nmc2 = splat_f32x2(-m_used * c);
        uint64_t la = 0, lb = 0;
        uint32_t pk[32];
#pragma unroll
        for (int k = 0; k < 32; ++k) {  // pair index
          const uint64_t x = fma_f32x2(pack_f32x2(r[2 * k], r[2 * k + 1]), c2, nmc2);
          uint64_t p;
          if ((POLY_MASK >> (k & 7)) & 1u) {
            p = exp2_poly_f32x2(x);
          } else {
            uint32_t xl, xh;
            unpack_f32x2(x, xl, xh);
            p = pack_f32x2(__float_as_uint(ex2_approx(__uint_as_float(xl))), __float_as_uint(ex2_approx(__uint_as_float(xh))));
          }
          if (k & 1) lb = add_f32x2(lb, p); else la = add_f32x2(la, p);
          uint32_t pl, ph;
          unpack_f32x2(p, pl, ph);
          pk[k] = pack_bf16x2(__uint_as_float(pl), __uint_as_float(ph));
        }
        tmem_st_32x32b_x32(tS, pk);  // P(j): 32 packed columns at the start of this S buffer
        uint32_t a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
      };
      if (kv_valid < FMHA2_KV) softmax_tile(std::true_type{});
      else softmax_tile(std::false_type{});
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[2 * t + b]);
    }
    // ---- final: O / l -> bf16 -> smem (this tile's Q buffer is free now) -> coalesced stores ----
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    const uint32_t stg = smem_u32(sQ + t * FMHA2_Q_BYTES);  // 128 rows x 256 B
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
        const int c16 = cc * 4 + g;
        sts_v4(stg + row * 256 + ((c16 ^ (row & 7)) << 4), v);
      }
    }
    __syncwarp();
    const int bb = bh / prm.H, h = bh - bb * prm.H;
    const int D = prm.H * 128;
    const int S_img = prm.S - prm.T;
#pragma unroll 1
    for (int it = 0; it < 16; ++it) {
      const int rr = q * 32 + it * 2 + (lane >> 4);
      const int c16 = lane & 15;
      const int pos = q_row0 + t * 128 + rr;
      if (pos < prm.S) {
        const uint4 v = lds_v4(stg + rr * 256 + ((c16 ^ (rr & 7)) << 4));
        bf16* dst = (pos < prm.T) ? prm.out_txt + ((size_t)bb * prm.T + pos) * D
                                  : prm.out_img + ((size_t)bb * S_img + (pos - prm.T)) * D;
        stg_v4(dst + h * 128 + c16 * 8, v);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace qimg
