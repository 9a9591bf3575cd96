// 2-CTA (cta_group::2) variant of the tcgen05 GEMM: a cluster of two CTAs on one TPC computes a
// 256 x 256 output tile with ONE tcgen05.mma.cta_group::2 (M = 256) per K step.  Each CTA stages its
// own 128 rows of A and only HALF of the B tile (128 of the 256 weight rows), so the shared-memory
// read traffic per SM drops from 96 B/clk (1-CTA, 128x256) to 64 B/clk and the pipeline holds 6 stages
// of 32 KB instead of 4 of 48 KB.  Same fused epilogues (epilogue_tile in qimg_gemm.cuh): each CTA
// drains its own 128 accumulator rows from its own TMEM.
//
// Synchronisation (leader = cluster rank 0):
//   full[s]   : leader's mbarrier only; both CTAs' TMA loads complete_tx on it (cp.async.bulk.tensor
//               .cta_group::2 with the leader's barrier address), the leader arms 2 x 32 KB
//   empty[s]  : one per CTA, released by tcgen05.commit.cta_group::2 ... multicast::cluster (mask 0b11)
//   tmem_full : one per CTA, same multicast commit after the last K block
//   tmem_empty: leader's, 16 arrivals (8 epilogue warps x 2 CTAs; the peer arrives remotely)
#pragma once

#include "qimg_gemm.cuh"

namespace qimg {

constexpr int GEMM2_STAGES = 6;
constexpr int GEMM2_GROUP_M = 16;  // default raster band: 16 x 256 rows = 4 bands at M = 16384 (8 -> 16: +1.5..2 % and half the weight re-reads, profiles/r02_gemm_group_m.log; 32 loses: the activation band falls out of L2) (qimg_set_gemm_group_m overrides; 1-CTA kernel: 2x as many 128-row tiles)
constexpr int GEMM2_A_BYTES = 128 * GEMM_BK * 2;
constexpr int GEMM2_B_BYTES = 128 * GEMM_BK * 2;
constexpr int GEMM2_STAGE_BYTES = GEMM2_A_BYTES + GEMM2_B_BYTES;
constexpr int GEMM2_SMEM_BYTES = GEMM2_STAGES * GEMM2_STAGE_BYTES + GEMM_EPI_STAGE_BYTES + 1024 + 256;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_cg2(uint32_t smem_dst, const CUtensorMap* m, uint32_t mbar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(m), "r"(mbar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_cg2(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_cg2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the mbarrier at this smem offset in BOTH CTAs of the pair when the issued MMAs complete
__device__ __forceinline__ void umma_commit_cg2_mc(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)), "h"(mask) : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_umma2_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                  const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                  const __grid_constant__ GemmParams prm) {
  constexpr int BN = 256;
  constexpr uint32_t IDESC = make_idesc_bf16(256, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + GEMM2_STAGES * GEMM2_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + GEMM_EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                            // [STAGES]  (leader's are the live ones)
  uint64_t* empty_bar = bars + GEMM2_STAGES;            // [STAGES]
  uint64_t* tmem_full = bars + 2 * GEMM2_STAGES;        // [2]
  uint64_t* tmem_empty = bars + 2 * GEMM2_STAGES + 2;   // [2]      (leader's are the live ones)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GEMM2_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  if (prm.skip && *prm.skip) return;  // uniform over the grid (both CTAs of every pair): nothing allocated or armed yet

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (prm.nprob > 1) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int i = 0; i < GEMM2_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 16);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_cg2<512>(tmem_ptr);
  tc_fence_before();
  cluster_sync_all();  // barriers of both CTAs initialised and visible cluster-wide, TMEM allocated
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; completion lands on the leader's full barrier) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = cluster_id; tile < prm.total_tiles; tile += num_clusters) {
      TileCoord tc = decode_tile(prm, tile, prm.group_m);
      const GemmProblem& P = prm.p[tc.pi];
      const CUtensorMap* ta = tc.pi ? &tmA1 : &tmA0;
      const CUtensorMap* tb = tc.pi ? &tmB1 : &tmB0;
      const int kblocks = (P.K + GEMM_BK - 1) / GEMM_BK;
      const int m_row = (tc.m_blk * 2 + (int)rank) * 128;
      const int n_row = tc.n_blk * BN + (int)rank * 128;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
          const uint32_t sb = sa + GEMM2_A_BYTES;
          const uint32_t leader_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * GEMM2_STAGE_BYTES);
          tma_load_2d_cg2(sa, ta, leader_full, kb * GEMM_BK, m_row);
          tma_load_2d_cg2(sb, tb, leader_full, kb * GEMM_BK, n_row);
        }
        __syncwarp();
        if (++stage == GEMM2_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; warp-uniform control flow, one elected lane issues) =====================
    if (rank == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < prm.total_tiles; tile += num_clusters) {
        TileCoord tc = decode_tile(prm, tile, prm.group_m);
        const GemmProblem& P = prm.p[tc.pi];
        const int kblocks = (P.K + GEMM_BK - 1) / GEMM_BK;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem + stage * GEMM2_STAGE_BYTES);
            const uint64_t adesc = make_kmajor_sw128_desc(sa);
            const uint64_t bdesc = make_kmajor_sw128_desc(sa + GEMM2_A_BYTES);
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k)
              umma_ss_cg2(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) != 0);
            umma_commit_cg2_mc(&empty_bar[stage]);
            if (kb == kblocks - 1) umma_commit_cg2_mc(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == GEMM2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue warps (both CTAs, own 128 rows) =====================
    constexpr int CHUNKS = BN / 64;
    constexpr int CH_PER_HALF = CHUNKS / 2;
    const int q = warp & 3;
    const int ew = warp - 2;
    const int chunk_lo = (ew >> 2) * CH_PER_HALF;
    const int chunk_hi = chunk_lo + CH_PER_HALF;
    const uint32_t stg = smem_u32(smem_epi + ew * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = cluster_id; tile < prm.total_tiles; tile += num_clusters) {
      TileCoord tc = decode_tile(prm, tile, prm.group_m);
      const GemmProblem& P = prm.p[tc.pi];
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      epilogue_tile<BN, EPI>(P, tc.m_blk * 2 + (int)rank, tc.n_blk, t_row, stg, lane, q, chunk_lo, chunk_hi);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[acc]), 0));
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  // the leader's MMAs read the peer's shared memory and both CTAs own half of the paired TMEM allocation:
  // nobody leaves before everyone is done
  tc_fence_before();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_cg2<512>(tmem_base);
  }
}

}  // namespace qimg
