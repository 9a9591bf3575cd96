// Joint attention, 128-row KV tiles, one softmax thread per query row, SHORT softmax->MMA chain (FMHA "v8").
//
// Cycle traces of the v4 kernel (tools/fmha_trace.py, profiles/r01_fmha_trace.md) show the KV-tile period is a serial
// chain per query tile:  QK -> [S visible] -> score load -> row max -> exponentials -> P stores -> [P visible] -> PV
// -> next QK, i.e. T = X + 2 MMA with X ~ 2250 cycles and MMA ~ 570; the tensor pipe idles ~1100 cycles per KV tile
// waiting for tile 0's P.  Two changes shorten X; TMA / MMA structure, TMEM map (S0|S1|O0|O1, P aliases S) and issue
// order are those of fmha_joint_kernel (qimg_fmha.cuh):
//   1. Delayed reference maximum.  Softmax is shift-invariant, so tile j is exponentiated against the reference the row
//      already has (the maximum over tiles < j, lazily updated) instead of first reducing its own maximum: the row-max
//      pass (~250 cycles) leaves the chain and becomes FMNMX3 work on the ALU pipe under the MUFU-bound exponentials.
//      The tile's maximum is only used afterwards, to decide (threshold 2^8, warp-uniform) whether O and l are rebased
//      before the next tile.  Only tile 0 reduces its maximum first.  P may exceed 1 (by the jump of the row maximum
//      inside one tile); exponents are clamped at 2^96 so that l and O stay finite in fp32 — a row whose scores
//      jump by more than 66 nats above everything seen before within a single tile loses the relative weights of the
//      clamped entries (never reached by RMS-normalised q/k; the exact pipelines 0-4 remain selectable).
//   2. P is handed to the MMA warp in four 32-column quarters (one mbarrier each): P*V of a quarter (2 UMMA k-steps)
//      runs while the next quarter is exponentiated, so only the last quarter's MMA stays on the chain.
#pragma once

#include <type_traits>

#include "qimg_fmha.cuh"

namespace qimg {

template <uint32_t POLY_MASK, bool PINGPONG>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_joint_kernel_v8(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
  uint64_t* p_ready = s_full + 2;        // [2 tiles][4 quarters]
  uint64_t* o_full = p_ready + 8;        // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int full_pairs = prm.S / 256 + ((prm.S % 256) > 128 ? 1 : 0);  // grid remap: see fmha_joint_kernel
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
  const bool two = q_row0 + 128 < prm.S;
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
      mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < 8; ++i) mbar_init(&p_ready[i], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer =====================
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
    // ===================== MMA issuer =====================
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
    // P*V of KV tile jj for query tile t, quarter by quarter as the softmax warpgroup releases P
    auto pv_quarters = [&](int t, int jj, long long& waited, bool tr) {
      const uint32_t va = smem_u32(sV + (jj % FMHA_VS) * FMHA_TILE_BYTES);
#pragma unroll
      for (int qd = 0; qd < 4; ++qd) {
        long long tt = 0;
        if (tr) tt = clock64();
        mbar_wait(&p_ready[t * 4 + qd], jj & 1);
        if (tr) waited += clock64() - tt;
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const int k = qd * 2 + kk;
            umma_ts(tO[t], tS[t] + k * 8, make_mnmajor_sw128_desc(va + k * 2048, 16384), IDESC_PV, (jj > 0 || k != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
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
      if (two && j > 0) pv_quarters(1, j - 1, w_p1, tr);
      if (elect_one()) {
        if (j > 0) umma_commit(&v_empty[(j - 1) % FMHA_VS]);  // V(j-1): PV0(j-1) and PV1(j-1) are both issued
        if (two) {
          issue_qk(1, ks);
          umma_commit(&s_full[1]);
        }
        umma_commit(&k_empty[ks]);
      }
      __syncwarp();
      if (tr) tt = clock64();
      mbar_wait(&v_full[j % FMHA_VS], (j / FMHA_VS) & 1);
      if (tr) w_v += clock64() - tt;
      pv_quarters(0, j, w_p0, tr);
    }
    if (two) pv_quarters(1, n_kv - 1, w_p1, tr);
    if (elect_one()) {
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
    // ===================== softmax / correction / output warps: one thread per query row =====================
    const int t = (warp - 2) >> 2;  // query tile handled by this warpgroup
    if (t == 0 || two) {
    const bool pingpong = PINGPONG && two;
    const int q = warp & 3;         // TMEM lane quarter
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128;
    const uint32_t tO = tmem_base + lane_off + 256 + t * 128;
    const float c = prm.scale_log2;
    float m_ref = 0.f;   // reference (raw score units) the exponentials of the current tile are taken against
    float m_next = 0.f;  // reference for the next tile (differs from m_ref when the row maximum jumped by > 2^8)
    float l = 0.f;       // running row sum, relative to m_ref
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
      auto softmax_tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        uint32_t r[128];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) tmem_ld_32x32b_x32(tS + cc * 32, r + cc * 32);
        // rebase O and l if the previous tile raised the row maximum by more than 2^8 (P*V(j-1) of this tile is
        // complete: s_full(j) was committed after it in the in-order tensor pipe); rare after the first tiles
        if (j > 0 && __any_sync(0xffffffffu, m_next != m_ref)) {
          const float f = ex2_approx((m_ref - m_next) * c);
          l *= f;
#pragma unroll 1
          for (int cc = 0; cc < 8; ++cc) {
            uint32_t o[16];
            tmem_ld_32x32b_x16(tO + cc * 16, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st_32x32b_x16(tO + cc * 16, o);
          }
          tmem_st_wait();
          m_ref = m_next;
        }
        tmem_ld_wait();
        if (tr) w_ld += clock64() - tt, tt = clock64();
        if (MASKED) {
#pragma unroll
          for (int i = 0; i < 128; ++i)
            if (i >= kv_valid) r[i] = 0xff800000u;  // -inf
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
        if (j == 0) {  // the first tile has no reference yet: reduce its maximum before exponentiating
#pragma unroll
          for (int i = 0; i < 128; i += 8) {
            mx0 = max3_f32(mx0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
            mx1 = max3_f32(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            mx2 = max3_f32(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
            mx3 = max3_f32(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
          }
          m_ref = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        }
        if (tr) w_mx += clock64() - tt, tt = clock64();
        if (pingpong) named_bar_sync(1 + t, 256);  // my turn on the XU pipe
        if (tr) w_pp += clock64() - tt, tt = clock64();
        const uint64_t c2 = splat_f32x2(c), nmc2 = splat_f32x2(-m_ref * c);
        uint64_t la = 0, lb = 0;
        // maximum of one 32-column quarter of the raw scores (4 independent FMNMX3 chains)
        auto quarter_max = [&](int qd) {
          float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            a0 = max3_f32(a0, __uint_as_float(r[qd * 32 + i]), __uint_as_float(r[qd * 32 + i + 1]));
            a1 = max3_f32(a1, __uint_as_float(r[qd * 32 + i + 2]), __uint_as_float(r[qd * 32 + i + 3]));
            a2 = max3_f32(a2, __uint_as_float(r[qd * 32 + i + 4]), __uint_as_float(r[qd * 32 + i + 5]));
            a3 = max3_f32(a3, __uint_as_float(r[qd * 32 + i + 6]), __uint_as_float(r[qd * 32 + i + 7]));
          }
          return fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
        };
        // exponentials of one quarter -> 16 packed bf16 pairs; CLAMP only on the (practically never taken) path where
        // the quarter's maximum exceeds the reference by more than 2^96
        auto exp_quarter = [&](int qd, uint32_t* pk, auto clamp_tag) {
          constexpr bool CLAMP = decltype(clamp_tag)::value;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = qd * 16 + i;  // pair index 0..63
            uint64_t x = fma_f32x2(pack_f32x2(r[2 * k], r[2 * k + 1]), c2, nmc2);
            if (CLAMP) {
              uint32_t xl, xh;
              unpack_f32x2(x, xl, xh);
              x = pack_f32x2(__float_as_uint(fminf(__uint_as_float(xl), 96.0f)), __float_as_uint(fminf(__uint_as_float(xh), 96.0f)));
            }
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
        };
        // Quarter qd+1's maximum is reduced while quarter qd is exponentiated (independent ALU-pipe work), so the
        // overflow guard is off the chain; quarter qd is released to the MMA warp only after quarter qd+1 has been
        // exponentiated, so the store-completion wait never stalls the exponentials (the last quarter waits once).
        float tile_max = (j == 0) ? m_ref : quarter_max(0);
        float qm = tile_max;
#pragma unroll
        for (int qd = 0; qd < 4; ++qd) {
          float qm_next = -INFINITY;
          if (j > 0 && qd < 3) qm_next = quarter_max(qd + 1);
          uint32_t pk[16];
          if (j > 0 && __any_sync(0xffffffffu, (qm - m_ref) * c > 96.0f)) exp_quarter(qd, pk, std::true_type{});
          else exp_quarter(qd, pk, std::false_type{});
          if (qd > 0) {  // release quarter qd-1: its stores were issued a whole quarter ago
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready[t * 4 + qd - 1]);
          }
          tmem_st_32x32b_x16(tS + qd * 16, pk);
          tile_max = fmaxf(tile_max, qm_next);
          qm = qm_next;
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t * 4 + 3]);
        if (tr) w_ex += clock64() - tt, tt = clock64();
        if (pingpong && !(t == 1 && j == n_kv - 1)) named_bar_arrive(1 + (t ^ 1), 256);  // hand the XU pipe over
        {
          uint32_t a0, a1, b0, b1;
          unpack_f32x2(la, a0, a1);
          unpack_f32x2(lb, b0, b1);
          l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
        }
        // reference for the next tile: lazily follow the row maximum
        m_next = m_ref;
        if (j > 0) {
          const float m_new = fmaxf(m_ref, tile_max);
          if ((m_new - m_ref) * c > 8.0f) m_next = m_new;
        }
      };
      if (kv_valid < 128) softmax_tile(std::true_type{});
      else softmax_tile(std::false_type{});
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
    // (a rebase still pending from the last tile would scale O and l alike and cancels in O / l)
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
