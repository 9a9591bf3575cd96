// Joint attention, third-generation pipeline (FMHA "v6"): fully decoupled softmax and tensor pipes.
//
// Measured problem of the earlier pipelines (profiles/r01_ncu_fmha_final.csv: tensor pipe 60 % active,
// softmax warps ~40 % of their samples waiting for S): per query tile the chain
//     S(j) ready -> softmax -> P(j) -> PV(j) -> QK(j+1) -> S(j+1) ready
// is serial because P(j) aliased S(j) in TMEM, so QK(j+1) could not be issued before PV(j).
// Here P has its OWN TMEM region and the KV tile is 80 rows so that everything fits in 512 columns:
//     S[t] 2 x 80 | P[t] 2 x 40 | (16 spare) | O[t] 2 x 128
// The softmax warpgroup releases S(j) as soon as the row is in registers (s_free), the MMA warp issues
// QK(j+1) immediately, and by the time exp/pack/store of P(j) is done S(j+1) is already waiting: the
// softmax warpgroups run back to back and the tensor pipe always has queued work.  The MMA warp no longer
// follows a fixed order; it polls the barriers and issues whichever of QK(t, .) / PV(t, .) is ready.
//
// Also new: CTAs whose second query tile lies completely beyond S (the ragged last pair: S = 4224 gives
// 16.5 pairs per head) skip that tile entirely and are scheduled LAST (blockIdx remap) so they fill the
// tail of the last wave: 1536 full + 96 half CTAs on 148 SMs finish in ~11.1 instead of 12 CTA-times.
#pragma once

#include <type_traits>

#include "qimg_fmha.cuh"

namespace qimg {

constexpr int FMHA3_KS = 3;
constexpr int FMHA3_VS = 3;
constexpr int FMHA3_Q_BYTES = 128 * 128 * 2;       // 32 KB per query tile (two 64-col SW128 slabs of 16 KB)
template <int KV>
constexpr int fmha3_smem_bytes() {
  return 2 * FMHA3_Q_BYTES + (FMHA3_KS + FMHA3_VS) * (2 * KV * 128) + 1024 + 512;
}

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_32x32b_x8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <uint32_t POLY_MASK, int FMHA3_KV>
__global__ void __launch_bounds__(FMHA_THREADS, 1)
fmha_joint_kernel_v6(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const __grid_constant__ FmhaParams prm) {
  static_assert(FMHA3_KV == 64 || FMHA3_KV == 80, "KV tile");
  constexpr int FMHA3_SLAB = FMHA3_KV * 128;      // one 64-column slab of a K/V tile (multiple of 1024 B)
  constexpr int FMHA3_KV_BYTES = 2 * FMHA3_SLAB;
  constexpr int P_BASE = 2 * FMHA3_KV;            // TMEM: S[t] at t*KV, P[t] at 2*KV + t*KV/2, O[t] at 256 + 128 t
  constexpr int P_COLS = FMHA3_KV / 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * FMHA3_Q_BYTES;
  uint8_t* sV = sK + FMHA3_KS * FMHA3_KV_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sV + FMHA3_VS * FMHA3_KV_BYTES);
  uint64_t* q_full = bars;                      // [1]
  uint64_t* k_full = bars + 1;                  // [KS]
  uint64_t* k_empty = k_full + FMHA3_KS;        // [KS]
  uint64_t* v_full = k_empty + FMHA3_KS;        // [VS]
  uint64_t* v_empty = v_full + FMHA3_VS;        // [VS]
  uint64_t* s_full = v_empty + FMHA3_VS;        // [2]  QK(t, j) complete
  uint64_t* s_free = s_full + 2;                // [2]  softmax(t, j) holds S(j) in registers
  uint64_t* p_ready = s_free + 2;               // [2]  P(t, j) stored
  uint64_t* pv_done = p_ready + 2;              // [2]  PV(t, j) complete
  uint64_t* o_full = pv_done + 2;               // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // work remap: query-tile pairs whose second tile is entirely out of range go LAST (see header)
  const int pairs = (prm.S + 255) / 256;
  const int full_pairs = prm.S / 256 + ((prm.S % 256) > 128 ? 1 : 0);   // pairs with both tiles (partly) valid
  const int n_bh = prm.B * prm.H;
  const int lin = blockIdx.x;
  int pair_idx, bh;
  if (lin < full_pairs * n_bh) {
    bh = lin / full_pairs;
    pair_idx = lin - bh * full_pairs;
  } else {
    const int r = lin - full_pairs * n_bh;      // half CTAs: one per (b,h) (pairs - full_pairs is 0 or 1)
    bh = r;
    pair_idx = full_pairs;
  }
  (void)pairs;
  const int q_row0 = pair_idx * 256;
  const int n_tiles = (q_row0 + 128 < prm.S) ? 2 : 1;
  const int n_kv = (prm.S + FMHA3_KV - 1) / FMHA3_KV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < FMHA3_KS; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < FMHA3_VS; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_ready[i], 4);
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
    // ===================== TMA producer =====================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, n_tiles * FMHA3_Q_BYTES);
      for (int t = 0; t < n_tiles; ++t)
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sQ + t * FMHA3_Q_BYTES + s * 16384, &tmQ, q_full, s * 64, q_row0 + t * 128, bh);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int ks = j % FMHA3_KS, vs = j % FMHA3_VS;
      mbar_wait(&k_empty[ks], ((j / FMHA3_KS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[ks], FMHA3_KV_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sK + ks * FMHA3_KV_BYTES + s * FMHA3_SLAB, &tmK, &k_full[ks], s * 64, j * FMHA3_KV, bh);
      }
      __syncwarp();
      mbar_wait(&v_empty[vs], ((j / FMHA3_VS) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[vs], FMHA3_KV_BYTES);
        for (int s = 0; s < 2; ++s)
          tma_load_3d(sV + vs * FMHA3_KV_BYTES + s * FMHA3_SLAB, &tmV, &v_full[vs], s * 64, j * FMHA3_KV, bh);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: polls readiness, issues whatever can run =====================
    constexpr uint32_t IDESC_QK = make_idesc_bf16(128, FMHA3_KV, 0, 0);
    constexpr uint32_t IDESC_PV = make_idesc_bf16(128, 128, 0, 1);  // B (=V) is MN-major
    auto issue_qk = [&](int t, int ks) {
      const uint32_t qa = smem_u32(sQ + t * FMHA3_Q_BYTES);
      const uint32_t ka = smem_u32(sK + ks * FMHA3_KV_BYTES);
      const uint32_t d = tmem_base + t * FMHA3_KV;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        umma_ss(d, make_kmajor_sw128_desc(qa + (k >> 2) * 16384 + (k & 3) * 32),
                make_kmajor_sw128_desc(ka + (k >> 2) * FMHA3_SLAB + (k & 3) * 32), IDESC_QK, k != 0);
      }
    };
    auto issue_pv = [&](int t, int vs, bool accumulate) {
      const uint32_t va = smem_u32(sV + vs * FMHA3_KV_BYTES);
      const uint32_t d = tmem_base + 256 + t * 128;
      const uint32_t p = tmem_base + P_BASE + t * P_COLS;
#pragma unroll
      for (int k = 0; k < FMHA3_KV / 16; ++k) {
        // A = P (bf16 pairs, 8 TMEM columns per K=16 step); B = V rows [16k,16k+16) x 128 (MN-major)
        umma_ts(d, p + k * 8, make_mnmajor_sw128_desc(va + k * 2048, FMHA3_SLAB), IDESC_PV, (accumulate || k != 0) ? 1u : 0u);
      }
    };
    auto ready = [&](uint64_t* bar, uint32_t parity) { return __all_sync(0xffffffffu, mbar_test(bar, parity)) != 0; };
    mbar_wait(q_full, 0);
    int jq[2] = {0, n_tiles == 2 ? 0 : n_kv};  // next QK / PV step per tile (an inactive tile is "finished")
    int jp[2] = {0, n_tiles == 2 ? 0 : n_kv};
    int k_rel = 0, v_rel = 0;                   // K / V tiles already handed back to the producer
    while (jp[0] < n_kv || jp[1] < n_kv) {
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        if (jq[t] < n_kv) {
          const int j = jq[t];
          if (ready(&k_full[j % FMHA3_KS], (j / FMHA3_KS) & 1) && (j == 0 || ready(&s_free[t], (j - 1) & 1))) {
            tc_fence_after();
            if (elect_one()) {
              issue_qk(t, j % FMHA3_KS);
              umma_commit(&s_full[t]);
            }
            __syncwarp();
            ++jq[t];
          }
        }
        if (jp[t] < jq[t]) {
          const int j = jp[t];
          if (ready(&p_ready[t], j & 1) && ready(&v_full[j % FMHA3_VS], (j / FMHA3_VS) & 1)) {
            tc_fence_after();
            if (elect_one()) {
              issue_pv(t, j % FMHA3_VS, j > 0);
              umma_commit(&pv_done[t]);
              if (j == n_kv - 1) umma_commit(&o_full[t]);
            }
            __syncwarp();
            ++jp[t];
          }
        }
      }
      // hand K / V tiles back once BOTH tiles have issued the MMAs that read them
      const int kmin = jq[0] < jq[1] ? jq[0] : jq[1];
      const int vmin = jp[0] < jp[1] ? jp[0] : jp[1];
      if (k_rel < kmin || v_rel < vmin) {
        if (elect_one()) {
          for (; k_rel < kmin; ++k_rel) umma_commit(&k_empty[k_rel % FMHA3_KS]);
          for (; v_rel < vmin; ++v_rel) umma_commit(&v_empty[v_rel % FMHA3_VS]);
        }
        __syncwarp();
        k_rel = kmin;
        v_rel = vmin;
      }
    }
  } else {
    // ===================== softmax / correction / output warps =====================
    const int t = (warp - 2) >> 2;  // query tile handled by this warpgroup
    if (t < n_tiles) {
      const int q = warp & 3;         // TMEM lane quarter
      const uint32_t lane_off = (uint32_t)(q * 32) << 16;
      const uint32_t tS = tmem_base + lane_off + t * FMHA3_KV;
      const uint32_t tP = tmem_base + lane_off + P_BASE + t * P_COLS;
      const uint32_t tO = tmem_base + lane_off + 256 + t * 128;
      const float c = prm.scale_log2;
      float m_used = -INFINITY;  // row max (raw score units) the exponentials are referenced to
      float l = 0.f;             // running row sum
      for (int j = 0; j < n_kv; ++j) {
        mbar_wait(&s_full[t], j & 1);
        tc_fence_after();
        uint32_t r[FMHA3_KV];
        tmem_ld_32x32b_x32(tS, r);
        tmem_ld_32x32b_x32(tS + 32, r + 32);
        if (FMHA3_KV > 64) tmem_ld_32x32b_x16(tS + 64, r + 64);
        tmem_ld_wait();
        // the score row is in registers: QK(j+1) may overwrite S right away
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);
        const int kv_valid = prm.S - j * FMHA3_KV;  // < 80 only on a ragged last tile
        auto softmax_tile = [&](auto masked_tag) {
          constexpr bool MASKED = decltype(masked_tag)::value;
          if (MASKED) {
#pragma unroll
            for (int i = 0; i < FMHA3_KV; ++i)
              if (i >= kv_valid) r[i] = 0xff800000u;  // -inf
          }
          float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < FMHA3_KV; i += 8) {
            mx0 = max3_f32(mx0, __uint_as_float(r[i]), __uint_as_float(r[i + 1]));
            mx1 = max3_f32(mx1, __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
            mx2 = max3_f32(mx2, __uint_as_float(r[i + 4]), __uint_as_float(r[i + 5]));
            mx3 = max3_f32(mx3, __uint_as_float(r[i + 6]), __uint_as_float(r[i + 7]));
          }
          const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
          bool pv_waited = (j == 0);
          if (j == 0) {
            m_used = mx;
          } else {
            const float m_new = fmaxf(m_used, mx);
            const bool need = (m_new - m_used) * c > 8.0f;
            if (__any_sync(0xffffffffu, need)) {
              // rare: rescale O and l.  PV(j-1) must be complete first.  pv_done[t] completes one phase per PV;
              // phases 0..j-2 are certainly complete here, so the parity of phase j-1 is unambiguous.
              mbar_wait(&pv_done[t], (j - 1) & 1);
              tc_fence_after();
              pv_waited = true;
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
          uint32_t pk[FMHA3_KV / 2];
#pragma unroll
          for (int k = 0; k < FMHA3_KV / 2; ++k) {  // pair index
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
          // P(j-1) must have been consumed by PV(j-1) before it is overwritten; the exponentials above ran
          // concurrently with that MMA, so this wait normally returns immediately.
          if (!pv_waited) {
            mbar_wait(&pv_done[t], (j - 1) & 1);
            tc_fence_after();
          }
          tmem_st_32x32b_x32(tP, pk);
          if (FMHA3_KV > 64) tmem_st_32x32b_x8(tP + 32, pk + 32);
          uint32_t a0, a1, b0, b1;
          unpack_f32x2(la, a0, a1);
          unpack_f32x2(lb, b0, b1);
          l += (__uint_as_float(a0) + __uint_as_float(a1)) + (__uint_as_float(b0) + __uint_as_float(b1));
        };
        if (kv_valid < FMHA3_KV) softmax_tile(std::true_type{});
        else softmax_tile(std::false_type{});
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_ready[t]);
      }
      // ---- final: O / l -> bf16 -> smem (this tile's Q buffer is free now) -> coalesced stores ----
      mbar_wait(&o_full[t], 0);
      tc_fence_after();
      const float inv_l = 1.0f / l;
      const uint32_t stg = smem_u32(sQ + t * FMHA3_Q_BYTES);  // 128 rows x 256 B
      const int row = q * 32 + lane;
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace qimg
