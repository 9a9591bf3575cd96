// Joint attention, FMHA "v10": ONE 128-row query tile per CTA, DOUBLE-BUFFERED scores, P in its own TMEM region.
//
// Why (profiles/r01_fmha_trace.md, r02 notes): in v4..v9 a CTA holds two query tiles whose S tiles fill TMEM together with
// the two O accumulators, P has to alias S, and therefore Q*K^T of KV tile j+1 cannot be issued before P*V of tile j has
// consumed P — per query tile the KV-tile period is the serial chain  QK -> softmax -> PV -> QK ...  (v9: ~3100 cycles
// per 256 rows where the tensor pipe needs 2048 and the XU pipe 2048; softmax warps idle ~45 % of the time waiting for S,
// the tensor pipe ~25 % waiting for P).  With one query tile per CTA the budget is
//     TMEM 512 columns = O (128) | S0 (128) | S1 (128) | P0 (64) | P1 (64)
// so the chain is broken twice: the scores live in REGISTERS (4 threads per row x 32 columns), the S buffer is handed
// back right after the TMEM load and QK(j+2) is issued while softmax(j) is still exponentiating; P(j) goes to its own
// buffer, so nothing waits for P*V either.  (v10a loaded the whole score tile before the first exponential: tcgen05.ld
// moves 64 B/clk per SM, 64 KB = 1024 cycles, serialised in front of 1024 cycles of XU work -> 2300 cycles per tile;
// v10b reads the scores in 16-column chunks under the exponentials.)  Steady state: softmax never waits (XU bound, 128x128 exponentials =
// 1024 cycles per KV tile at 16/clk/SM), the tensor pipe runs PV(j), QK(j+2) back to back (2 x 512 cycles) in its shadow.
// K/V are fetched once per 128 (not 256) query rows: 2x the L2->SM traffic of v9, i.e. ~30 % of the measured L2 peak
// (v9: 4.7 TB/s = 15 %, profiles/r01_ncu_fmha_v9.csv) — bandwidth that was idle.
//
// Roles (576 threads): warp 0 TMA producer (Q once, K through a 3-stage and V through a 2-stage mbarrier ring), warp 1 TMEM
// allocator + tcgen05.mma issuer, warps 2-17 softmax: thread = (row, column quarter).  Softmax is the v9 scheme: tile 0
// reduces its row maximum first; tiles j >= 1 are exponentiated against the running reference maximum, the tile's own
// maximum (exchanged between the four threads of a row through shared memory) only decides the lazy rebase of O / l at
// the start of the next tile (threshold 2^8; the rebase waits for P*V(j-1) explicitly, and only then).  Arguments beyond
// 2^100 raise prm.overflow (see qimg_fmha.cuh).
#pragma once

#include <type_traits>

#include "qimg_fmha.cuh"

namespace qimg {

constexpr int FMHA7_THREADS = 32 * (2 + 16);
constexpr int FMHA7_KS = 3;
constexpr int FMHA7_VS = 2;
constexpr int FMHA7_SMEM_BYTES = (1 + FMHA7_KS + FMHA7_VS) * FMHA_TILE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 4096 /*xch*/;

template <uint32_t POLY_MASK>
__global__ void __launch_bounds__(FMHA7_THREADS, 1)
fmha_joint_kernel_v10(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FmhaParams prm) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 1 tile (reused as output staging)
  uint8_t* sK = smem + FMHA_TILE_BYTES;                 // KS tiles
  uint8_t* sV = sK + FMHA7_KS * FMHA_TILE_BYTES;        // VS tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FMHA7_VS * FMHA_TILE_BYTES);
  uint64_t* q_full = bars;                  // [1]
  uint64_t* k_full = bars + 1;              // [KS]
  uint64_t* k_empty = k_full + FMHA7_KS;
  uint64_t* v_full = k_empty + FMHA7_KS;    // [VS]
  uint64_t* v_empty = v_full + FMHA7_VS;
  uint64_t* s_full = v_empty + FMHA7_VS;    // [2]  QK(j) complete
  uint64_t* s_empty = s_full + 2;           // [2]  all 16 softmax warps hold the scores of the buffer in registers
  uint64_t* p_ready = s_empty + 2;          // [2]  P(j) stored by all 16 warps
  uint64_t* p_free = p_ready + 2;           // [2]  P*V(j) complete -> the P buffer may be overwritten
  uint64_t* o_done = p_free + 2;            // [1]  one phase per P*V (rebase waits on it)
  uint64_t* o_full = o_done + 1;            // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 1);
  float* xch = reinterpret_cast<float*>(bars + 32);  // [parity][column quarter][128 rows] partial row maxima (4 KB)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (prm.skip && *prm.skip) return;  // uniform over the grid: nothing allocated or armed yet
  const int q_tiles = (prm.S + 127) / 128;
  const int bh = blockIdx.x / q_tiles;            // head-major: the 33 CTAs of a head stream the same K/V through L2
  const int q_row0 = (blockIdx.x - bh * q_tiles) * 128;
  const int n_kv = q_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA7_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FMHA7_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 16);
      mbar_init(&p_ready[i], 16);
      mbar_init(&p_free[i], 1);
    }
    mbar_init(o_done, 1);
    mbar_init(o_full, 1);
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
      mbar_arrive_expect_tx(q_full, FMHA_TILE_BYTES);
      for (int s = 0; s < 2; ++s) tma_load_3d(sQ + s * 16384, &tmQ, q_full, s * 64, q_row0, bh);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int ks = j % FMHA7_KS, vs = j % FMHA7_VS;
      mbar_wait(&k_empty[ks], ((j / FMHA7_KS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[ks], FMHA_TILE_BYTES);
        for (int s = 0; s < 2; ++s) tma_load_3d(sK + ks * FMHA_TILE_BYTES + s * 16384, &tmK, &k_full[ks], s * 64, j * 128, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[vs], ((j / FMHA7_VS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[vs], FMHA_TILE_BYTES);
        for (int s = 0; s < 2; ++s) tma_load_3d(sV + vs * FMHA_TILE_BYTES + s * 16384, &tmV, &v_full[vs], s * 64, j * 128, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);  // B (=V) is MN-major
    const uint32_t tO = tmem_base;
    const uint32_t qa = smem_u32(sQ);
    auto issue_qk = [&](int j) {  // S_{j&1} = Q K(j)^T
      const uint32_t ka = smem_u32(sK + (j % FMHA7_KS) * FMHA_TILE_BYTES);
      const uint32_t tS = tmem_base + 128 + (j & 1) * 128;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tS, make_kmajor_sw128_desc(qa + off), make_kmajor_sw128_desc(ka + off), IDESC_QK, k != 0);
      }
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
      issue_qk(0);
      umma_commit(&s_full[0]);
      umma_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) {  // scores of the NEXT tile: needs K(j+1) and the S buffer softmax(j-1) has emptied
        const int jn = j + 1;
        mbar_wait(&k_full[jn % FMHA7_KS], (jn / FMHA7_KS) & 1);
        mbar_wait(&s_empty[jn & 1], ((jn >> 1) & 1) ^ 1);
        tc_fence_after();
        if (elect_one()) {
          issue_qk(jn);
          umma_commit(&s_full[jn & 1]);
          umma_commit(&k_empty[jn % FMHA7_KS]);
        }
        __syncwarp();
      }
      mbar_wait(&v_full[j % FMHA7_VS], (j / FMHA7_VS) & 1);
      mbar_wait(&p_ready[j & 1], (j >> 1) & 1);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t va = smem_u32(sV + (j % FMHA7_VS) * FMHA_TILE_BYTES);
        const uint32_t tP = tmem_base + 384 + (j & 1) * 64;
#pragma unroll
        for (int k = 0; k < 8; ++k)  // A = P (bf16 pairs, 8 TMEM columns per K=16 step); B = V rows [16k, 16k+16) x 128 (MN-major)
          umma_ts(tO, tP + k * 8, make_mnmajor_sw128_desc(va + k * 2048, 16384), IDESC_PV, (j > 0 || k != 0) ? 1u : 0u);
        umma_commit(&v_empty[j % FMHA7_VS]);
        umma_commit(&p_free[j & 1]);
        umma_commit(o_done);
        if (j == n_kv - 1) umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ===================== softmax / correction / output warps: FOUR threads per query row =====================
    // warp -> (TMEM lane quarter q = warp & 3 — the hardware's rule —, column quarter cq): the four warps of a lane quarter
    // own the same 32 rows and 32 score columns each; they share named barrier 1 + q (128 threads).
    const int cq = (warp - 2) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int grp_bar = 1 + q;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t tS0 = tmem_base + lane_off + 128 + cq * 32;   // + (j & 1) * 128
    const uint32_t tP0 = tmem_base + lane_off + 384 + cq * 16;   // + (j & 1) * 64
    const uint32_t tO = tmem_base + lane_off + cq * 32;
    const float c = prm.scale_log2;
    float m_ref = 0.f;              // reference (raw score units) the exponentials of the current tile are taken against
    float l = 0.f;                  // partial row sum over my column quarter, relative to m_ref
    float my_tile_max = -INFINITY;  // maximum of my 32 columns of the previous tile
    for (int j = 0; j < n_kv; ++j) {
      const int b = j & 1;
      mbar_wait(&s_full[b], (j >> 1) & 1);
      tc_fence_after();
      const uint32_t tS = tS0 + b * 128;
      const int kv_valid = prm.S - j * 128 - cq * 32;  // valid columns of my quarter (< 32 only on a ragged last tile)
      float* my_x = xch + ((j & 1) * 4 + cq) * 128 + row;
      const float* grp_x = xch + (j & 1) * 4 * 128 + row;
      auto softmax_tile = [&](auto masked_tag) {
        constexpr bool MASKED = decltype(masked_tag)::value;
        // tcgen05.ld moves 64 B/clk per SM: the 64 KB score tile takes ~1024 cycles to read, as long as its exponentials
        // take on the XU pipe, so the two must overlap — 16 columns at a time, the next load in flight (the first version
        // of this kernel loaded the whole tile first and ran at 2300 cycles per tile)
        uint32_t ra[16], rb[16];
        tmem_ld_32x32b_x16(tS, ra);
        float tile_max;
        if (j == 0) {
          // ---- first tile: no reference yet -> the row maximum of this tile first (exchange between the four threads) ----
          tmem_ld_32x32b_x16(tS + 16, rb);
          tmem_ld_wait();
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              if (i >= kv_valid) ra[i] = 0xff800000u;
              if (16 + i >= kv_valid) rb[i] = 0xff800000u;
            }
          }
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; i += 2) {
            mx0 = max3_f32(mx0, __uint_as_float(ra[i]), __uint_as_float(ra[i + 1]));
            mx1 = max3_f32(mx1, __uint_as_float(rb[i]), __uint_as_float(rb[i + 1]));
          }
          tile_max = fmaxf(mx0, mx1);
          *my_x = tile_max;
          named_bar_sync(grp_bar, 128);
          m_ref = fmaxf(fmaxf(grp_x[0], grp_x[128]), fmaxf(grp_x[256], grp_x[384]));
        } else {
          // ---- later tiles: agree on the PREVIOUS tile's row maximum; rebase O / l if it rose by more than 2^8 ----
          *my_x = my_tile_max;
          named_bar_sync(grp_bar, 128);
          const float m_new = fmaxf(m_ref, fmaxf(fmaxf(grp_x[0], grp_x[128]), fmaxf(grp_x[256], grp_x[384])));
          const bool need = (m_new - m_ref) * c > 8.0f;  // identical in the four warps of the group (same rows, same maxima)
          if (__any_sync(0xffffffffu, need)) {
            mbar_wait(o_done, (j - 1) & 1);  // P*V(j-1) accumulated under the old reference must be complete
            tc_fence_after();
            const float f = need ? ex2_approx((m_ref - m_new) * c) : 1.0f;
            l *= f;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
              uint32_t o[16];
              tmem_ld_32x32b_x16(tO + cc * 16, o);
              tmem_ld_wait();  // (also completes the score chunk load issued above)
#pragma unroll
              for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
              tmem_st_32x32b_x16(tO + cc * 16, o);
            }
            tmem_st_wait();
            if (need) m_ref = m_new;
          }
        }
        // ---- exp2((s - m_ref) c) -> bf16 P, two chunks of 16 columns; the partial row sum stays in fp32 ----
        const uint64_t c2 = splat_f32x2(c), nmc2 = splat_f32x2(-m_ref * c);
        uint64_t la = 0, lb = 0;
        float mx0 = -INFINITY, mx1 = -INFINITY;
        uint32_t pk[16];
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t* cur = ch ? rb : ra;
          if (j > 0) {
            tmem_ld_wait();
            if (ch == 0) tmem_ld_32x32b_x16(tS + 16, rb);
            if (MASKED) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (ch * 16 + i >= kv_valid) cur[i] = 0xff800000u;
            }
          }
          if (ch == 1) {  // all my score columns are in registers: QK(j+2) may overwrite this S buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[b]);
          }
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            if (j > 0) {
              if (kk & 1) mx1 = max3_f32(mx1, __uint_as_float(cur[2 * kk]), __uint_as_float(cur[2 * kk + 1]));
              else mx0 = max3_f32(mx0, __uint_as_float(cur[2 * kk]), __uint_as_float(cur[2 * kk + 1]));
            }
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
            pk[ch * 8 + kk] = pack_bf16x2(__uint_as_float(pl), __uint_as_float(ph));
          }
        }
        uint32_t a0, a1, b0, b1;
        unpack_f32x2(la, a0, a1);
        unpack_f32x2(lb, b0, b1);
        l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
        if (j > 0) tile_max = fmaxf(mx0, mx1);
        my_tile_max = (j == 0) ? m_ref : tile_max;
        // guard (see qimg_fmha.cuh): exact while no argument exceeded 2^FMHA_OVF_LOG2, otherwise flag the launch
        if (j > 0 && (tile_max - m_ref) * c > FMHA_OVF_LOG2) *prm.overflow = 1;
        mbar_wait(&p_free[b], ((j >> 1) & 1) ^ 1);  // P*V(j-2) has consumed this P buffer (long ago, normally)
        tmem_st_32x32b_x16(tP0 + b * 64, pk);
      };
      if (kv_valid < 32) softmax_tile(std::true_type{});
      else softmax_tile(std::false_type{});
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);
    }
    // ---- final: combine the four partial row sums, O / l -> bf16 -> smem -> coalesced stores ----
    float* my_l = xch + cq * 128 + row;
    const float* grp_l = xch + row;
    named_bar_sync(grp_bar, 128);  // the maxima of the last tile have been read by everyone
    *my_l = l;
    named_bar_sync(grp_bar, 128);
    const float inv_l = 1.0f / ((grp_l[0] + grp_l[128]) + (grp_l[256] + grp_l[384]));
    mbar_wait(o_full, 0);
    tc_fence_after();
    const uint32_t stg = smem_u32(sQ);  // 128 rows x 256 B (Q is no longer needed: every Q K^T has completed)
    {
      uint32_t o[32];
      tmem_ld_32x32b_x32(tO, o);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 v;
        v.x = pack_bf16x2(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
        v.y = pack_bf16x2(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
        v.z = pack_bf16x2(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
        v.w = pack_bf16x2(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
        const int c16 = cq * 4 + g;  // 16-byte chunk index within the 256 B row
        sts_v4(stg + row * 256 + ((c16 ^ (row & 7)) << 4), v);
      }
    }
    named_bar_sync(grp_bar, 128);  // all four quarters of my 32 rows are staged
    const int bb = bh / prm.H, h = bh - bb * prm.H;
    const int D = prm.H * 128;
    const int S_img = prm.S - prm.T;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {  // this warp stores 8 of the group's 32 rows
      const int rr = q * 32 + cq * 8 + it * 2 + (lane >> 4);
      const int c16 = lane & 15;
      const int pos = q_row0 + rr;
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
