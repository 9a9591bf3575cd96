// tcgen05 (UMMA) persistent, warp-specialised bf16 GEMM for sm_100a with fused epilogues.
//
//   out[M,N] = A[M,K] * W[N,K]^T  (nn.Linear layout: both operands K-major) + epilogue
//
// Replaces (reference file:line, all cuBLAS + separate ATen kernels there):
//   to_qkv / add_kv_proj + norm_q/k + rope + cat  qwen_image_transformer.py:380-416   -> EPI_QKV
//   to_out[0] / to_add_out + gate*x + residual    :452-456, 586-587                   -> EPI_BIAS_GATE_RES
//   img_mlp/txt_mlp net.0.proj + gelu(tanh)       :491,501,591,596                    -> EPI_BIAS_GELU
//   net.2 + gate*x + residual                     :592,597                            -> EPI_BIAS_GATE_RES
//   img_in / txt_in / proj_out                    :743,759,798                        -> EPI_BIAS
//   tensor-parallel to_out / net.2 (K sharded): fp32 partial sums pushed to the row owner's peer buffer -> EPI_PARTIAL_F32
// The image and text streams (different weights, M_txt << M_img) are GROUPED in one
// launch so the small text GEMM fills the tail instead of starving 140 SMs.
//
// Structure (one CTA per SM, 320 threads):
//   warp 0   : TMA producer  (cp.async.bulk.tensor 2D, SWIZZLE_128B, 4-stage mbarrier ring)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (128 x BN x 16 per instruction)
//   warps 2-9: epilogue, two warps per scheduler (warps 2-5 take the first BN/2 columns, 6-9 the
//              second): tcgen05.ld accumulator rows -> math (packed bf16x2 where the reference rounds
//              per op) -> swizzled smem staging -> 128-byte coalesced global stores (residual / gate
//              fused there)
//   two TMEM accumulator stages (2 x BN columns) so the epilogue of tile i overlaps the
//   main loop of tile i+1.
#pragma once

#include "qimg_common.cuh"

namespace qimg {

enum GemmEpilogue { EPI_BIAS = 0, EPI_BIAS_GELU = 1, EPI_BIAS_GATE_RES = 2, EPI_QKV = 3, EPI_PARTIAL_F32 = 4 };

struct GemmProblem {
  int M, N, K;
  int rows_per_batch;  // rows of this stream per image (S_img or T)
  const bf16* bias;    // [N]
  bf16* out;           // EPI_BIAS / EPI_BIAS_GELU: output [M, ldo]; EPI_BIAS_GATE_RES: residual stream x (in/out)
  int ldo;
  const bf16* gate;  // EPI_BIAS_GATE_RES: gate[b * gate_stride + n]
  long long gate_stride;
  // EPI_QKV
  bf16* q;  // joint [B, H, S_joint, 128]
  bf16* k;
  bf16* v;
  const bf16* nq_w;  // RMSNorm weights [128]
  const bf16* nk_w;
  const bf16* cos;  // [rows_per_batch, 64]
  const bf16* sin;
  int S_joint, pos_off, H;
  float eps;
  // EPI_PARTIAL_F32 (tensor-parallel row-parallel linear): fp32 partial sums are PUSHED to the receive buffer of the rank
  // that owns the row (peer memory over NVLink).  Rows are split contiguously and balanced over tp_size owners; owner o
  // keeps the rows of source rank s at tp_recv[o] + ((s * tp_recv_rows + tp_recv_row_off + local_row) * N) floats.
  float* tp_recv[8];
  int tp_size, tp_rank, tp_recv_rows, tp_recv_row_off;
  // Sequence parallelism (rows of A are a rank's OWN slice of the stream): row_base = global index of local row 0 (batch /
  // position / gate lookups use global rows).  EPI_QKV with sp_size > 1: head h goes to the rank that owns it,
  // sp_q/k/v[h / H_local] with the local head index h % H_local (H_local = H / sp_size) — the all-to-all in front of
  // Ulysses attention (reference attention/parallel/ulysses.py:110-112), fused into the GEMM epilogue as peer stores.
  int row_base;
  int sp_size;
  bf16* sp_q[8];
  bf16* sp_k[8];
  bf16* sp_v[8];
  // tile bookkeeping (filled by the host launcher)
  int m_tiles, n_tiles, tile_begin;
};

struct GemmParams {
  GemmProblem p[2];
  int nprob;
  int total_tiles;
  int group_m;      // raster band height in tiles (launcher: qimg_set_gemm_group_m)
  const int* skip;  // optional device predicate: non-zero -> the kernel exits at once (step-cache reuse, qimg_tea_decide)
};

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 320;
constexpr int GEMM_GROUP_M = 16;
constexpr int GEMM_EPI_STAGE_BYTES = 8 * 32 * 128;  // 8 warps x 32 rows x 128 B

template <int BN>
constexpr int gemm_smem_bytes() {
  return GEMM_STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + GEMM_EPI_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
}

struct TileCoord {
  int pi, m_blk, n_blk;
};

__device__ __forceinline__ TileCoord decode_tile(const GemmParams& prm, int tile, int group_m = GEMM_GROUP_M) {
  TileCoord tc;
  tc.pi = (prm.nprob > 1 && tile >= prm.p[1].tile_begin) ? 1 : 0;
  const GemmProblem& P = prm.p[tc.pi];
  int t = tile - P.tile_begin;
  int per_band = group_m * P.n_tiles;
  int band = t / per_band;
  int within = t - band * per_band;
  int m0 = band * group_m;
  int gm = min(group_m, P.m_tiles - m0);
  tc.m_blk = m0 + within % gm;
  tc.n_blk = within / gm;
  return tc;
}

// One accumulator tile (128 rows x BN columns, rows [m_blk*128, +128)) -> global memory.  Executed by the 8
// epilogue warps of a CTA; warp (q = lane quarter, chunk range) owns 32 rows x (chunk_hi-chunk_lo)*64 columns.
// EPI_PARTIAL_F32: the compute + collective fusion of the tensor-parallel row-parallel linears.  The accumulator tile
// leaves TMEM as fp32 and goes straight over NVLink into the receive buffer of the rank that owns each row (the
// reduce-scatter half of the all-reduce), tile by tile, under the main loop of the next tile (two TMEM accumulator
// stages) — no bf16 rounding of partial sums, no local round trip through HBM, no separate reduction-input pass.
template <int BN>
__device__ __forceinline__ void epilogue_tile_partial(const GemmProblem& P, int m_blk, int n_blk, uint32_t t_row, uint32_t stg,
                                                      int lane, int q, int chunk_lo, int chunk_hi) {
  // destination row pointers of this lane's 8 store rows (gm0 + 4 it), computed once per tile: owner o of a row and its
  // index within o's slice advance incrementally (one division per tile, none in the store loop)
  const int gm0 = m_blk * GEMM_BM + q * 32 + (lane >> 3);
  const int base = P.M / P.tp_size, extra = P.M % P.tp_size;
  float* dst_row[8];
  {
    int o, start;
    const int cut = extra * (base + 1);
    if (gm0 < cut) {
      o = gm0 / (base + 1);
      start = o * (base + 1);
    } else {
      o = base > 0 ? extra + (gm0 - cut) / base : P.tp_size - 1;
      start = cut + (o - extra) * base;
    }
    int size = base + (o < extra ? 1 : 0);
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int gm = gm0 + it * 4;
      while (gm >= start + size && o + 1 < P.tp_size) {
        start += size;
        ++o;
        size = base + (o < extra ? 1 : 0);
      }
      dst_row[it] = P.tp_recv[o] + ((size_t)P.tp_rank * P.tp_recv_rows + P.tp_recv_row_off + (gm - start)) * (size_t)P.N;
    }
  }
#pragma unroll 1
  for (int chunk = chunk_lo; chunk < chunk_hi; ++chunk) {
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const int n0 = n_blk * BN + chunk * 64 + half * 32;
      uint32_t r[32];
      tmem_ld_32x32b_x32(t_row + chunk * 64 + half * 32, r);
      tmem_ld_wait();
      // this thread's row: 32 fp32 = 128 B -> staging (16 B chunk index XOR row&7), then 128 B-per-row coalesced stores
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(r[j * 4], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]));
      __syncwarp();
      const int c16 = lane & 7;
      const int gn = n0 + c16 * 4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + (lane >> 3);
        const uint4 v = lds_v4(stg + rr * 128 + ((c16 ^ (rr & 7)) << 4));
        if (gm0 + it * 4 < P.M && gn < P.N) stg_v4(dst_row[it] + gn, v);
      }
      __syncwarp();
    }
  }
}

template <int BN, int EPI>
__device__ __forceinline__ void epilogue_tile(const GemmProblem& P, int m_blk, int n_blk, uint32_t t_row, uint32_t stg,
                                              int lane, int q, int chunk_lo, int chunk_hi) {
  if (EPI == EPI_PARTIAL_F32) {
    epilogue_tile_partial<BN>(P, m_blk, n_blk, t_row, stg, lane, q, chunk_lo, chunk_hi);
    return;
  }
  const int m_own = m_blk * GEMM_BM + q * 32 + lane;  // this thread's accumulator row

  // batch / in-batch index of the first row of this warp's 32-row slab: one division per tile, rows
  // then advance incrementally (no per-row integer division in the store loops)
  const int slab_m0 = m_blk * GEMM_BM + q * 32;  // local row (addresses); + row_base = global row (batch, position)
  const int slab_b0 = (P.row_base + slab_m0) / P.rows_per_batch;
  const int slab_i0 = (P.row_base + slab_m0) - slab_b0 * P.rows_per_batch;
  float rstd = 0.f;
  const bf16* cos_row = nullptr;
  const bf16* sin_row = nullptr;
  if (EPI == EPI_QKV) {
    int i_own = slab_i0 + lane;
    while (i_own >= P.rows_per_batch) i_own -= P.rows_per_batch;
    if (m_own >= P.M) i_own = 0;
    cos_row = P.cos + (size_t)i_own * 64;
    sin_row = P.sin + (size_t)i_own * 64;
  }

#pragma unroll 1
  for (int chunk = chunk_lo; chunk < chunk_hi; ++chunk) {
    const int n0 = n_blk * BN + chunk * 64;
    int which = 0, head = 0, half = 0;
    if (EPI == EPI_QKV) {
      const int D = P.N / 3;
      which = n0 / D;                // 0 = q, 1 = k, 2 = v
      head = (n0 - which * D) >> 7;  // head index
      half = (n0 >> 6) & 1;          // which 64-column half of the head
      if (which < 2 && half == 0 && n0 < P.N) {
        // pass 1 over the whole head (128 columns): sum of squares of bf16(acc + bias)
        float ss0 = 0.f, ss1 = 0.f;
#pragma unroll 1
        for (int c4 = 0; c4 < 4; ++c4) {
          uint32_t t[32];
          tmem_ld_32x32b_x32(t_row + chunk * 64 + c4 * 32, t);
          const uint4* bp = reinterpret_cast<const uint4*>(P.bias + n0 + c4 * 32);
          uint4 bv[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) bv[j] = __ldg(bp + j);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t bw[4] = {bv[j].x, bv[j].y, bv[j].z, bv[j].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const uint32_t v = pack_bf16x2(__uint_as_float(t[j * 8 + e * 2]) + bf16lo(bw[e]),
                                             __uint_as_float(t[j * 8 + e * 2 + 1]) + bf16hi(bw[e]));
              ss0 = fmaf(bf16lo(v), bf16lo(v), ss0);
              ss1 = fmaf(bf16hi(v), bf16hi(v), ss1);
            }
          }
        }
        rstd = rsqrtf((ss0 + ss1) * (1.0f / 128.0f) + P.eps);
      }
    }
    uint32_t r[64];
    tmem_ld_32x32b_x32(t_row + chunk * 64, r);
    tmem_ld_32x32b_x32(t_row + chunk * 64 + 32, r + 32);
    // this row's cos/sin (per-thread addresses) are fetched while the TMEM read is in flight;
    // bias / norm weights are warp-uniform broadcast loads issued just in time (keeps registers < 168)
    uint4 cv[4], sv[4];
    if (EPI == EPI_QKV && which < 2 && n0 < P.N) {
      const uint4* cp = reinterpret_cast<const uint4*>(cos_row + half * 32);
      const uint4* sp = reinterpret_cast<const uint4*>(sin_row + half * 32);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        cv[j] = __ldg(cp + j);
        sv[j] = __ldg(sp + j);
      }
    }
    tmem_ld_wait();

    // ---- math on this thread's 64 columns -> 32 packed bf16x2 words ----
    uint32_t pk[32];
    if (n0 < P.N) {
      const uint4* bp = reinterpret_cast<const uint4*>(P.bias + n0);
      const uint4* np = reinterpret_cast<const uint4*>((which == 0 ? P.nq_w : P.nk_w) + half * 64);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const uint4 bvj = __ldg(bp + j);
        const uint32_t bw[4] = {bvj.x, bvj.y, bvj.z, bvj.w};
        uint4 nvj = make_uint4(0, 0, 0, 0);
        if (EPI == EPI_QKV && which < 2) nvj = __ldg(np + j);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // v = bf16(acc + bias): the Linear's bf16 output
          uint32_t v = pack_bf16x2(__uint_as_float(r[j * 8 + e * 2]) + bf16lo(bw[e]),
                                   __uint_as_float(r[j * 8 + e * 2 + 1]) + bf16hi(bw[e]));
          if (EPI == EPI_BIAS_GELU) {
            v = pack_bf16x2(gelu_tanh_fast(bf16lo(v)), gelu_tanh_fast(bf16hi(v)));
          }
          if (EPI == EPI_QKV && which < 2) {
            const uint32_t nw4[4] = {nvj.x, nvj.y, nvj.z, nvj.w};
            // RMSNorm: fp32 normalise -> bf16 -> * weight -> bf16   (vLLM rms_norm)
            uint32_t x = bmul2(pack_bf16x2(bf16lo(v) * rstd, bf16hi(v) * rstd), nw4[e]);
            // interleaved RoPE, every op rounded to bf16: x*cos + rotate_half(x)*sin with
            // rotate_half(x) = (-x[2i+1], x[2i]); pair index within this 64-col chunk = j*4 + e
            const int pi = j * 4 + e;  // 0..31 -> 16-byte vector pi/8, word (pi%8)/2, half pi&1
            const uint4 c4 = cv[pi >> 3], s4 = sv[pi >> 3];
            const uint32_t cw4[4] = {c4.x, c4.y, c4.z, c4.w}, sw4[4] = {s4.x, s4.y, s4.z, s4.w};
            const uint32_t cw = cw4[(pi & 7) >> 1], sw = sw4[(pi & 7) >> 1];
            const uint32_t cc = (pi & 1) ? dup_hi(cw) : dup_lo(cw);
            const uint32_t ssn = (pi & 1) ? dup_hi(sw) : dup_lo(sw);
            const uint32_t rot = swap_halves(x) ^ 0x00008000u;  // (-x_hi, x_lo)
            v = badd2(bmul2(x, cc), bmul2(rot, ssn));
          }
          pk[j * 4 + e] = v;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) pk[j] = 0;
    }

    // ---- stage to smem (16 B chunk index XOR row&7 -> conflict-free), then coalesced stores ----
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(pk[j * 4], pk[j * 4 + 1], pk[j * 4 + 2], pk[j * 4 + 3]));
    }
    __syncwarp();
    const int c16 = lane & 7;
    const int gn = n0 + c16 * 8;
    const int gm0 = m_blk * GEMM_BM + q * 32 + (lane >> 3);
    uint4 yv[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int rr = it * 4 + (lane >> 3);
      yv[it] = lds_v4(stg + rr * 128 + ((c16 ^ (rr & 7)) << 4));
    }
    if (EPI == EPI_BIAS || EPI == EPI_BIAS_GELU) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int gm = gm0 + it * 4;
        if (gm < P.M && gn < P.N) stg_v4(P.out + (size_t)gm * P.ldo + gn, yv[it]);
      }
    } else if (EPI == EPI_BIAS_GATE_RES) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {  // two batches of 4 rows: loads first, then math + stores
        uint4 xv[4], gv[4];
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int gm = gm0 + (hh * 4 + i4) * 4;
          if (gm < P.M && gn < P.N) {
            xv[i4] = ldg_v4(P.out + (size_t)gm * P.ldo + gn);
            int b = slab_b0, i = slab_i0 + (gm - slab_m0);
            while (i >= P.rows_per_batch) {
              i -= P.rows_per_batch;
              ++b;
            }
            gv[i4] = __ldg(reinterpret_cast<const uint4*>(P.gate + (size_t)b * P.gate_stride + gn));
          }
        }
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int it = hh * 4 + i4;
          const int gm = gm0 + it * 4;
          if (gm < P.M && gn < P.N) {
            // x = bf16(x + bf16(gate * y))
            uint4 o;
            o.x = badd2(xv[i4].x, bmul2(gv[i4].x, yv[it].x));
            o.y = badd2(xv[i4].y, bmul2(gv[i4].y, yv[it].y));
            o.z = badd2(xv[i4].z, bmul2(gv[i4].z, yv[it].z));
            o.w = badd2(xv[i4].w, bmul2(gv[i4].w, yv[it].w));
            stg_v4(P.out + (size_t)gm * P.ldo + gn, o);
          }
        }
      }
    } else {  // EPI_QKV: scatter into the joint [B,H,S,128] head-major layout
      bf16* base = (which == 0) ? P.q : (which == 1 ? P.k : P.v);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int gm = gm0 + it * 4;
        if (gm < P.M && gn < P.N) {
          int b = slab_b0, i = slab_i0 + (gm - slab_m0);
          while (i >= P.rows_per_batch) {
            i -= P.rows_per_batch;
            ++b;
          }
          if (P.sp_size > 1) {  // sequence parallel: the head's owner holds heads [o * Hl, (o + 1) * Hl) as local heads
            const int Hl = P.H / P.sp_size, o = head / Hl;
            bf16* pb = (which == 0) ? P.sp_q[o] : (which == 1 ? P.sp_k[o] : P.sp_v[o]);
            const size_t off = (((size_t)b * Hl + (head - o * Hl)) * P.S_joint + (P.pos_off + i)) * 128 + half * 64 + c16 * 8;
            stg_v4(pb + off, yv[it]);
          } else {
            const size_t off = (((size_t)b * P.H + head) * P.S_joint + (P.pos_off + i)) * 128 + half * 64 + c16 * 8;
            stg_v4(base + off, yv[it]);
          }
        }
      }
    }
    __syncwarp();
  }
}

template <int BN, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmB0,
                 const __grid_constant__ CUtensorMap tmA1, const __grid_constant__ CUtensorMap tmB1,
                 const __grid_constant__ GemmParams prm) {
  static_assert(BN == 64 || BN == 128 || BN == 256, "BN");
  constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
  constexpr int B_BYTES = BN * GEMM_BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;  // two accumulator stages
  constexpr uint32_t IDESC = make_idesc_bf16(GEMM_BM, BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + GEMM_STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + GEMM_EPI_STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [STAGES]
  uint64_t* empty_bar = bars + GEMM_STAGES;        // [STAGES]
  uint64_t* tmem_full = bars + 2 * GEMM_STAGES;    // [2]
  uint64_t* tmem_empty = bars + 2 * GEMM_STAGES + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (prm.skip && *prm.skip) return;  // uniform over the grid: nothing allocated or armed yet

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA0);
    tma_prefetch_desc(&tmB0);
    if (prm.nprob > 1) {
      tma_prefetch_desc(&tmA1);
      tma_prefetch_desc(&tmB1);
    }
    for (int i = 0; i < GEMM_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===================== TMA producer (warp-uniform control flow, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      TileCoord tc = decode_tile(prm, tile, prm.group_m);
      const GemmProblem& P = prm.p[tc.pi];
      const CUtensorMap* ta = tc.pi ? &tmA1 : &tmA0;
      const CUtensorMap* tb = tc.pi ? &tmB1 : &tmB0;
      const int kblocks = (P.K + GEMM_BK - 1) / GEMM_BK;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(sa, ta, &full_bar[stage], kb * GEMM_BK, tc.m_blk * GEMM_BM);
          tma_load_2d(sb, tb, &full_bar[stage], kb * GEMM_BK, tc.n_blk * BN);
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (warp-uniform control flow, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
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
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + A_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BK / 16; ++k) {
            // +32 B per K=16 step inside the 128 B swizzle row (descriptor start field is >>4)
            umma_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);                       // frees the smem slot when these MMAs complete
          if (kb == kblocks - 1) umma_commit(&tmem_full[acc]);  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == GEMM_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps =====================
    constexpr int CHUNKS = BN / 64;                   // 64-column chunks per tile
    constexpr int CH_PER_HALF = (CHUNKS + 1) / 2;     // BN=256: each warpgroup owns one 128-column head
    const int q = warp & 3;                           // TMEM lane quarter this warp may access
    const int ew = warp - 2;                          // staging slot 0..7
    const int chunk_lo = (ew >> 2) * CH_PER_HALF;
    const int chunk_hi = (chunk_lo + CH_PER_HALF < CHUNKS) ? chunk_lo + CH_PER_HALF : CHUNKS;
    const uint32_t stg = smem_u32(smem_epi + ew * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < prm.total_tiles; tile += gridDim.x) {
      TileCoord tc = decode_tile(prm, tile, prm.group_m);
      const GemmProblem& P = prm.p[tc.pi];
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
      epilogue_tile<BN, EPI>(P, tc.m_blk, tc.n_blk, t_row, stg, lane, q, chunk_lo, chunk_hi);
      // all TMEM reads of this accumulator stage are complete -> hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

}  // namespace qimg
