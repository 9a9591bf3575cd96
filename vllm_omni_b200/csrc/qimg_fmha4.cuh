// Joint attention, 128-row KV tiles, TWO softmax threads per query row (FMHA "v7").
// Same TMA / MMA structure, TMEM map (S0|S1|O0|O1, P aliases S) and issue order as fmha_joint_kernel
// (qimg_fmha.cuh); only the softmax side differs: 16 softmax warps, each thread owns 64 of the 128 score columns
// of its row.  Motivation (profiles/r01_ncu_fmha_final.csv and the v4 source-level samples): with one thread per
// row a softmax step took ~2100 cycles, ~1000 of them XU-serial (128 MUFU.EX2 x 8 cycles per warp) and ~1000
// latency-bound FMNMX/F2FP/TMEM work that a single warp per scheduler cannot overlap with its own MUFUs.
#pragma once

#include <type_traits>

#include "qimg_fmha.cuh"

namespace qimg {

constexpr int FMHA4_THREADS = 32 * (2 + 16);
constexpr int FMHA4_SMEM_BYTES = FMHA_SMEM_BYTES + 4096;

template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(FMHA4_THREADS, 1)
fmha_joint_kernel_v7(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
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
  float* xch = reinterpret_cast<float*>(o_full + 4);  // [tile][parity][half][128 rows] partial row maxima (4 KB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (prm.skip && *prm.skip) return;  // uniform over the grid: nothing allocated or armed yet
  const int n_bh = prm.B * prm.H;
  int bh, q_row0;
  bool two;
  if (prm.single_tile) {  // one query tile per CTA (the tile-1 softmax warps idle): small grids only
    const int q_tiles = (prm.S + 127) / 128;
    bh = blockIdx.x / q_tiles;
    q_row0 = (blockIdx.x - bh * q_tiles) * 128;
    two = false;
  } else {
    const FmhaWork work = fmha_decode_cta(blockIdx.x, prm.S, n_bh);  // head-major, half pairs lagged (qimg_fmha.cuh)
    bh = work.bh;
    q_row0 = work.pair_idx * 256;
    two = q_row0 + 128 < prm.S;  // is the second query tile (partly) in range?
  }
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
      mbar_init(&p_ready[i], 8);
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
        // (P sits 32 columns into the S tile: see the softmax warps' aliasing note)
        umma_ts(tO[t], tS[t] + 32 + k * 8, make_mnmajor_sw128_desc(va + k * 2048, 16384), IDESC_PV,
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
    // ===================== softmax / correction / output warps: TWO threads per query row =====================
    // 16 warps: tile t = (warp-2)/8, column half hh = ((warp-2)/4)&1 (score columns [64 hh, 64 hh + 64)), TMEM lane
    // quarter q = warp & 3.  The two warps of a (tile, quarter) pair sit on the same scheduler, so one can run its
    // FMNMX/F2FP/TMEM phases while the other keeps the XU pipe busy; each thread's dependent chains are half as long.
    // The partial row maxima are exchanged through shared memory (one 64-thread named barrier per KV tile); partial
    // row sums stay private until the end.
    const int t = (warp - 2) >> 3;
    if (t == 0 || two) {
    const int hh = ((warp - 2) >> 2) & 1;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int pair_bar = 1 + t * 4 + q;  // named barrier of this (tile, quarter) warp pair, 64 threads
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS = tmem_base + lane_off + t * 128 + hh * 64;   // my 64 score columns
    // P (64 packed columns) aliases S columns [32, 96): the hh=0 thread's 32 P columns land on ITS OWN score columns
    // [32, 64), the hh=1 thread's on its own [64, 96).  Pass 2 re-reads the scores chunk by chunk while P chunks are being
    // stored, so each thread walks its chunks in the order that only ever overwrites scores it has already consumed:
    // hh=0 descending (chunk c -> P columns 32+8c, inside chunk >= c), hh=1 ascending (P columns 64+8c, inside chunk <= c).
    const uint32_t tP = tmem_base + lane_off + t * 128 + 32 + hh * 32;   // my 32 packed P columns
    const uint32_t tO = tmem_base + lane_off + 256 + t * 128 + hh * 64;
    const float c = prm.scale_log2;
    float m_used = -INFINITY;
    float l = 0.f;  // partial row sum over my column half
    const bool tr = kFmhaTrace && prm.trace != nullptr && blockIdx.x == 200 && hh == 0 && q == 0;
    long long w_s = 0, w_ld = 0, w_x = 0, w_pp = 0, w_ex = 0, w_tl = 0, tt = 0;
    const long long t_begin = kFmhaTrace ? clock64() : 0;
    for (int j = 0; j < n_kv; ++j) {
      if (tr) tt = clock64();
      mbar_wait(&s_full[t], j & 1);
      if (tr) w_s += clock64() - tt, tt = clock64();
      tc_fence_after();
      const int kv_valid = prm.S - j * 128 - hh * 64;  // valid columns of my half (< 64 only on a ragged last tile)
      float* my_x = xch + ((t * 2 + (j & 1)) * 2 + hh) * 128 + row;
      const float* other_x = xch + ((t * 2 + (j & 1)) * 2 + (hh ^ 1)) * 128 + row;
      auto softmax_tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // ---- pass 1: partial row maximum over my 64 score columns (the scores are NOT kept in registers: with 18 warps
        // the per-thread budget is 96 registers, and holding 64 scores + 16 packed P words across the exp phase spilled
        // ~100 words per thread per tile; TMEM reads are cheap, so pass 2 reads the scores again in 16-column chunks) ----
        float part;
        {
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
          part = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
        }
        *my_x = part;
        uint32_t ra[16], rb[16];
        tmem_ld_32x32b_x16(tS + (hh ? 0 : 48), ra);  // first chunk of pass 2: its latency hides behind the exchange barrier
        if (tr) w_ld += clock64() - tt, tt = clock64();
        named_bar_sync(pair_bar, 64);  // partner's partial maximum is visible (buffers alternate with j)
        if (tr) w_x += clock64() - tt, tt = clock64();
        const float mx = fmaxf(part, *other_x);
        if (j == 0) {
          m_used = mx;
        } else {
          const float m_new = fmaxf(m_used, mx);
          const bool need = (m_new - m_used) * c > 8.0f;  // identical in both warps of the pair (same rows, same mx)
          if (__any_sync(0xffffffffu, need)) {
            const float f = ex2_approx((m_used - m_new) * c);
            l *= f;
#pragma unroll 1
            for (int cc = 0; cc < 4; ++cc) {  // my half of the O columns
              uint32_t o[16];
              tmem_ld_32x32b_x16(tO + cc * 16, o);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_32x32b_x16(tO + cc * 16, o);
            }
            tmem_st_wait();
            m_used = m_new;
          }
        }
        if (tr) w_pp += clock64() - tt, tt = clock64();
        const uint64_t c2 = splat_f32x2(c), nmc2 = splat_f32x2(-m_used * c);
        uint64_t la = 0, lb = 0;
        // ---- pass 2: exp2((s - m) c) -> bf16 P, 16 columns at a time, next chunk's TMEM load in flight ----
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t* cur = (ch & 1) ? rb : ra;
          uint32_t* nxt = (ch & 1) ? ra : rb;
          const int cidx = hh ? ch : 3 - ch;  // chunk of 16 score columns handled in this round
          tmem_ld_wait();
          if (ch < 3) tmem_ld_32x32b_x16(tS + (hh ? cidx + 1 : cidx - 1) * 16, nxt);
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (cidx * 16 + i >= kv_valid) cur[i] = 0xff800000u;
          }
          uint32_t pk[8];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint64_t x = fma_f32x2(pack_f32x2(cur[2 * kk], cur[2 * kk + 1]), c2, nmc2);
            uint64_t p;
            if ((POLY_MASK >> kk) & 1u) {
              p = exp2_poly_f32x2(x);
            } else {
              uint32_t xl, xh;
              unpack_f32x2(x, xl, xh);
              p = pack_f32x2(__float_as_uint(ex2_approx(__uint_as_float(xl))), __float_as_uint(ex2_approx(__uint_as_float(xh))));
            }
            if (kk & 1) lb = add_f32x2(lb, p); else la = add_f32x2(la, p);
            uint32_t pl, ph;
            unpack_f32x2(p, pl, ph);
            pk[kk] = pack_bf16x2(__uint_as_float(pl), __uint_as_float(ph));
          }
          tmem_st_32x32b_x8(tP + cidx * 8, pk);
        }
        if (tr) w_ex += clock64() - tt, tt = clock64();
        uint32_t a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
      };
      if (kv_valid < 64) softmax_tile(std::true_type{});
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
      o[2] = w_ld;   // pass 1: score loads + partial row maximum
      o[3] = w_x;    // pair barrier (max exchange)
      o[4] = w_pp;   // rescale check + ping-pong barrier
      o[5] = w_ex;   // pass 2: exponentials, P stores
      o[6] = w_tl;   // store wait, fence, arrive
    }
    // ---- final: combine the two partial row sums, O / l -> bf16 -> smem -> coalesced stores ----
    float* my_l = xch + ((t * 2 + 0) * 2 + hh) * 128 + row;           // xch is idle now (last use: max of tile n_kv-1,
    const float* other_l = xch + ((t * 2 + 0) * 2 + (hh ^ 1)) * 128 + row;  // ordered by the sync below for parity 1 too)
    named_bar_sync(pair_bar, 64);
    *my_l = l;
    named_bar_sync(pair_bar, 64);
    const float inv_l = 1.0f / (l + *other_l);
    mbar_wait(&o_full[t], 0);
    tc_fence_after();
    const uint32_t stg = smem_u32(sQ + t * FMHA_TILE_BYTES);  // 128 rows x 256 B
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO + cc * 32, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
        const int c16 = hh * 8 + cc * 4 + g;  // 16-byte chunk index within the 256 B row
        sts_v4(stg + row * 256 + ((c16 ^ (row & 7)) << 4), v);
      }
    }
    named_bar_sync(pair_bar, 64);  // both halves of my 32 rows are staged
    const int b = bh / prm.H, h = bh - b * prm.H;
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {  // this warp stores 16 of the pair's 32 rows
      const int rr = q * 32 + hh * 16 + it * 2 + (lane >> 4);
      const int c16 = lane & 15;
      const int pos = q_row0 + t * 128 + rr;
      if (pos < prm.S) {
        const uint4 v = lds_v4(stg + rr * 256 + ((c16 ^ (rr & 7)) << 4));
        stg_v4(fmha_out_ptr(prm, b, h, pos, c16), v);  // local buffer, or the row owner's over NVLink (sequence parallel)
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
