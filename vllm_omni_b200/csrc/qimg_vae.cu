// VAE kernels for sm_100a: the post-step of QwenImagePipeline.forward — AutoencoderKLQwenImage._decode
// (pipeline_qwen_image.py:736-747 -> autoencoder_kl_qwenimage.py:839-862) — and the edit pipelines' pre-step, _encode
// (pipeline_qwen_image_edit.py:458-480 -> :793-812), both for single-frame inputs.
//
// The reference runs these in fp32 with cuDNN (TF32 tensor-core convolutions by default) over NCHW tensors, ~70 separate ATen
// / cuDNN launches with an fp32 pad + Conv3d on a three-frame tensor per layer.  Here:
//   * activations live in HBM as fp32 NHWC ([image, y, x, channel]); a 3x3 "same" convolution is an implicit GEMM
//       out[pixel, co] = sum_{tap, ci} x[pixel + tap, ci] * w[co, tap, ci]
//     on the 5th-gen tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).  The activation operand is a 4-D TMA
//     box {32 channels, x, y, 1 image} loaded at the tap's shifted coordinates — the zero padding of the convolution is
//     TMA's out-of-bounds fill: no padded copy, no im2col buffer.  Two kernels:
//       conv2_tf32_kernel (default)  one box {32, 8 x, 18 y} per (horizontal tap, channel block) serves the three vertical
//                                    taps as shifted descriptor views; 16 x 16 pixel patch x 96 / 128 / 192 channels per CTA
//       conv_tf32_kernel  (first)    one box {32, 16 x, 8 y} per tap, 128 pixels x 128 channels per CTA; also the stride-2
//                                    resamplers of the encoder (the tensor map walks x and y with element stride 2)
//     both warp-specialised (TMA warp, MMA warp, 8 epilogue warps), SWIZZLE_128B, mbarrier rings, TMEM accumulator stages
//     so that the epilogue (bias + residual add in fp32, coalesced NHWC stores) of tile i overlaps the main loop of tile i+1.
//     1x1 convolutions and the attention GEMMs of the mid blocks are the same kernels with taps = 1.
//   * TF32 inputs / fp32 accumulation is the arithmetic the reference's own GPU path uses for these layers
//     (torch.backends.cudnn.allow_tf32 defaults to True), so parity against the fp32 oracle is at the reference's own level.
//   * the bandwidth-bound pieces are single-pass row kernels: RMS-norm (+ SiLU), nearest x2 upsample, row softmax, the
//     16 -> 16 post_quant_conv fused with the NCHW -> NHWC layout change, and conv_out (96 -> 3 channels: FMA-pipe direct
//     convolution fused with the clamp, the NHWC -> NCHW change and the uint8 post-process; N = 3 is no tensor-core shape).
// What bounds the convolution kernels (shared-memory port; measurements): DESIGN.md §3 "VAE decode".
#include "../../include/qimg_b200.h"

#include <cstdlib>
#include <cstring>

#include "qimg_common.cuh"
#include "qimg_host.cuh"

namespace qimg {

// ---------------------------------------------------------------------------------------------------------------------
// tcgen05 kind::tf32 + 4-D TMA
// ---------------------------------------------------------------------------------------------------------------------
// instruction descriptor: [4,6) c_format = 1 (F32), [7,10) a_format = 2 (TF32), [10,13) b_format = 2, K-major A and B,
// [17,23) N >> 3, [24,29) M >> 4   (cute/arch/mma_sm100_desc.hpp, F16F32Format::TF32 = 2)
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem], 128 x N x 8 per instruction
__device__ __forceinline__ void umma_ss_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}

constexpr int CONV_BM = 128;       // 16 x 8 pixel patch
constexpr int CONV_PX = 16;
constexpr int CONV_PY = 8;
constexpr int CONV_BN = 128;       // output channels per tile
constexpr int CONV_BK = 32;        // fp32 elements per K block = one 128-byte swizzle row
constexpr int CONV_STAGES = 5;
constexpr int CONV_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int CONV_A_BYTES = CONV_BM * CONV_BK * 4;
constexpr int CONV_B_BYTES = CONV_BN * CONV_BK * 4;
constexpr int CONV_STAGE_BYTES = CONV_A_BYTES + CONV_B_BYTES;
constexpr int CONV_EPI_BYTES = 8 * 32 * 128;
constexpr int CONV_SMEM_BYTES = CONV_STAGES * CONV_STAGE_BYTES + CONV_EPI_BYTES + 1024 + 256;

struct ConvParams {
  int N, H, W;       // images, rows, columns (output = input geometry: stride 1, "same" padding)
  int cin_blocks;    // ceil(Cin / 32)
  int taps;          // 9 = 3x3, 1 = 1x1
  int stride;        // 1, or 2 (v1 kernel only: 3x3 over an input zero-padded by one row / column at the bottom / right)
  int Cout;
  int ldo, ldr;      // pixel strides (floats) of out / res
  int tiles_x, tiles_y, n_tiles, total_tiles;
  const float* bias; // [Cout] or nullptr
  const float* res;  // [N, H, W, ldr] or nullptr: out = conv + bias + res
  float* out;        // [N, H, W, ldo]
};

struct ConvTile {
  int n, x0, y0, n_blk;
};
__device__ __forceinline__ ConvTile conv_decode_tile(const ConvParams& P, int tile) {
  ConvTile t;
  const int m_tile = tile / P.n_tiles;  // the n tiles of one pixel patch are adjacent: they run concurrently and share A in L2
  t.n_blk = tile - m_tile * P.n_tiles;
  const int per_img = P.tiles_x * P.tiles_y;
  t.n = m_tile / per_img;
  const int rem = m_tile - t.n * per_img;
  const int ty = rem / P.tiles_x;
  t.y0 = ty * CONV_PY;
  t.x0 = (rem - ty * P.tiles_x) * CONV_PX;
  return t;
}

__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvParams P) {
  constexpr int TMEM_COLS = 2 * CONV_BN;
  constexpr uint32_t IDESC = make_idesc_tf32(CONV_BM, CONV_BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_epi = smem + CONV_STAGES * CONV_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + CONV_EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + CONV_STAGES;
  uint64_t* tmem_full = bars + 2 * CONV_STAGES;
  uint64_t* tmem_empty = bars + 2 * CONV_STAGES + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * CONV_STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < CONV_STAGES; ++i) {
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
  const int kblocks = P.taps * P.cin_blocks;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const ConvTile t = conv_decode_tile(P, tile);
      int tap = 0, cb = 0;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sa = smem + stage * CONV_STAGE_BYTES;
          uint8_t* sb = sa + CONV_A_BYTES;
          // stride 1: taps at -1..1 ("same" padding); stride 2: taps at 0..2 of the input pixel (2x, 2y) — the reference's
          // ZeroPad2d((0, 1, 0, 1)) + Conv2d(3, stride 2) — the tensor map then walks x and y with element stride 2
          const int off = (P.taps == 9 && P.stride == 1) ? -1 : 0;
          const int dy = P.taps == 9 ? tap / 3 + off : 0;
          const int dx = P.taps == 9 ? tap - (tap / 3) * 3 + off : 0;
          mbar_arrive_expect_tx(&full_bar[stage], CONV_STAGE_BYTES);
          // rows of the box: channel fastest, then x, then y -> 128 rows of 128 B = the K-major SWIZZLE_128B A tile;
          // coordinates outside the image (negative or >= W / H) are zero-filled by TMA: the convolution's padding
          tma_load_4d(sa, &tmA, &full_bar[stage], cb * CONV_BK, t.x0 * P.stride + dx, t.y0 * P.stride + dy, t.n);
          tma_load_2d(sb, &tmB, &full_bar[stage], kb * CONV_BK, t.n_blk * CONV_BN);
        }
        __syncwarp();
        if (++cb == P.cin_blocks) {
          cb = 0;
          ++tap;
        }
        if (++stage == CONV_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * CONV_BN;
      for (int kb = 0; kb < kblocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = smem_u32(smem + stage * CONV_STAGE_BYTES);
          const uint64_t adesc = make_kmajor_sw128_desc(sa);
          const uint64_t bdesc = make_kmajor_sw128_desc(sa + CONV_A_BYTES);
#pragma unroll
          for (int k = 0; k < CONV_BK / 8; ++k)  // K = 8 tf32 = 32 B per instruction inside the 128 B swizzle row
            umma_ss_tf32(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == kblocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == CONV_STAGES) {
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
    // ===================== epilogue warps: warp (q, half) owns accumulator rows q*32..+32, columns half*64..+64 =========
    const int q = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;
    const uint32_t stg = smem_u32(smem_epi + ew * 4096);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const ConvTile t = conv_decode_tile(P, tile);
      // the 8 rows this lane stores (tile-local row q*32 + it*4 + lane/8): pixel index, -1 when outside the image
      long long pix[8];
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rl = q * 32 + it * 4 + (lane >> 3);
        const int x = t.x0 + (rl & (CONV_PX - 1)), y = t.y0 + (rl >> 4);
        pix[it] = (x < P.W && y < P.H) ? ((long long)t.n * P.H + y) * P.W + x : -1;
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + acc * CONV_BN;
#pragma unroll 1
      for (int step = 0; step < 2; ++step) {
        const int col = half * 64 + step * 32;
        const int n0 = t.n_blk * CONV_BN + col;
        if (n0 >= P.Cout) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_row + col, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(r[j * 4], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]));
        __syncwarp();
        const int c16 = lane & 7;
        const int gn = n0 + c16 * 4;
        if (gn < P.Cout) {
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias) bv = __ldg(reinterpret_cast<const float4*>(P.bias + gn));
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + (lane >> 3);
            const uint4 v = lds_v4(stg + rr * 128 + ((c16 ^ (rr & 7)) << 4));
            if (pix[it] >= 0) {
              float4 o = make_float4(__uint_as_float(v.x) + bv.x, __uint_as_float(v.y) + bv.y, __uint_as_float(v.z) + bv.z,
                                     __uint_as_float(v.w) + bv.w);
              if (P.res) {
                const float4 rv = *reinterpret_cast<const float4*>(P.res + (size_t)pix[it] * P.ldr + gn);
                o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
              }
              *reinterpret_cast<float4*>(P.out + (size_t)pix[it] * P.ldo + gn) = o;
            }
          }
        }
        __syncwarp();
      }
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

// ---------------------------------------------------------------------------------------------------------------------
// v2: the same implicit GEMM with 2.3x fewer operand bytes per MAC.  v1 above is bound by L2 -> shared-memory operand
// delivery (~14.5 TB/s at every layer shape, profiles/r02_vae_ncu_summary.md): every tap re-fetches the activation patch and
// every 128-pixel tile re-streams the weights.  Here
//   * ONE TMA box {32 channels, 8 x, 18 y} per (horizontal tap, channel block) serves all THREE vertical taps: rows of the
//     box are ordered x fastest, so every image row is one 8-row SWIZZLE_128B group (1024 B) and the view shifted by dy is
//     simply the K-major descriptor started (dy + 1) * 1024 B further — no copy, no extra load;
//   * a CTA owns MT = 2 such 8 x 16 pixel blocks (a 16 x 16 patch, two TMEM accumulators per stage) that share every weight
//     tile in shared memory, so weights are streamed once per 256 pixels.
//   * the tile is BN = 96 / 128 / 192 output channels wide — the decoder's layers are 96, 192 and 384 wide, so no MMA column
//     is padding (at BN = 128 a 96-wide layer wastes a quarter and a 192-wide layer a third of the tensor-pipe time).
//   Rings: weight tiles in an NB-slot mbarrier ring (5 x 16 KB, or 4 x 24 KB at BN = 192), one slot per (tap, channel block)
//   unit; activation boxes in an NA-slot ring (3, or 2 at BN = 192) filled together with the first unit of their group
//   (dy = -1) and released implicitly: unit u waits for the completion of unit u - NB before its slot is refilled, and
//   tcgen05.commit orders all earlier MMAs, so when group g's boxes are loaded every MMA of group g - NA has retired
//   (its last unit 3 (g - NA) + 2 <= 3 g - NB  <=>  NB <= 3 NA - 2).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int C2_NB_MAX = 7;
constexpr int C2_BOX9_BYTES = 18 * 8 * CONV_BK * 4;    // 18 KB: {32 ch, 8 x, 18 y}
constexpr int C2_BOX1_BYTES = 16 * 8 * CONV_BK * 4;    // 16 KB: {32 ch, 8 x, 16 y}
// Ring sizes by tile width (weight slots of BN x 128 B).  The main loop is bound by TMA round-trip latency x ring depth
// (~2-3 k cycles under load, profiles/r02_vae_ncu_summary.md), so every configuration fills the 227 KB: BN = 96 -> 7 x 12 KB
// weight slots + 3 activation slots, BN = 128 -> 5 x 16 KB + 3, BN = 192 -> 4 x 24 KB + 2.  NB <= 3 NA - 2 keeps the
// implicit release of the activation slots valid (see above).
__host__ __device__ constexpr int c2_b_slot(int BN) { return BN * CONV_BK * 4; }
__host__ __device__ constexpr int c2_nb(int BN) { return BN > 128 ? 4 : (BN > 96 ? 5 : 7); }
__host__ __device__ constexpr int c2_na(int BN) { return BN > 128 ? 2 : 3; }
__host__ __device__ constexpr int c2_a_region(int BN) { return c2_na(BN) * 2 * C2_BOX9_BYTES; }
__host__ __device__ constexpr int c2_smem_bytes(int BN) {
  return c2_a_region(BN) + c2_nb(BN) * c2_b_slot(BN) + CONV_EPI_BYTES + 1024 + 256;
}
static_assert(c2_smem_bytes(96) <= 232448, "shared memory");
static_assert(c2_nb(96) <= 3 * c2_na(96) - 2 && c2_nb(128) <= 3 * c2_na(128) - 2 && c2_nb(192) <= 3 * c2_na(192) - 2, "rings");
static_assert(c2_smem_bytes(128) <= 232448 && c2_smem_bytes(192) <= 232448, "shared memory");

struct Conv2Tile {
  int n, x0, y0, n_blk;
};
template <int MT>
__device__ __forceinline__ Conv2Tile conv2_decode_tile(const ConvParams& P, int tile) {
  Conv2Tile t;
  const int m_tile = tile / P.n_tiles;
  t.n_blk = tile - m_tile * P.n_tiles;
  const int per_img = P.tiles_x * P.tiles_y;
  t.n = m_tile / per_img;
  const int rem = m_tile - t.n * per_img;
  const int ty = rem / P.tiles_x;
  t.y0 = ty * 16;
  t.x0 = (rem - ty * P.tiles_x) * (8 * MT);
  return t;
}

// BN = output channels per tile = N of the MMA (96 / 128 / 192: the decoder's widths are 96, 192, 384 — no padded columns);
// ACC = TMEM accumulator stages (2 = the epilogue of tile i overlaps the main loop of tile i + 1; MT * BN * ACC <= 512).
template <int MT, int BN, int ACC>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv2_tf32_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvParams P) {
  static_assert(BN % 32 == 0 && BN <= 192 && MT * BN * ACC <= 512, "tile");
  constexpr int TMEM_COLS = (MT * BN * ACC <= 256) ? 256 : 512;
  constexpr uint32_t IDESC = make_idesc_tf32(CONV_BM, BN);
  constexpr int B_TILE_BYTES = BN * CONV_BK * 4;
  constexpr int B_SLOT = c2_b_slot(BN), NB9 = c2_nb(BN), NA9 = c2_na(BN), A_REGION = c2_a_region(BN);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + A_REGION;
  uint8_t* smem_epi = smem_b + NB9 * B_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + CONV_EPI_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C2_NB_MAX;
  uint64_t* tmem_full = bars + 2 * C2_NB_MAX;
  uint64_t* tmem_empty = bars + 2 * C2_NB_MAX + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * C2_NB_MAX + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < C2_NB_MAX; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < ACC; ++i) {
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

  // ring geometry (see the header comment): 3x3 -> groups of G = 3 units sharing one activation slot; 1x1 -> G = 1
  const bool k3 = P.taps == 9;
  const int G = k3 ? 3 : 1;
  const int a_box = k3 ? C2_BOX9_BYTES : C2_BOX1_BYTES;
  const int a_slot = MT * a_box;
  const int na1 = A_REGION / a_slot < NB9 ? A_REGION / a_slot : NB9;
  const int NA = k3 ? NA9 : na1;
  const int NB = k3 ? NB9 : na1;  // G = 1: a unit's activation slot is free once unit u - NB retired -> NB <= NA
  const int ngroups = G * P.cin_blocks;  // 3x3: (dx, channel block); 1x1: channel block

  if (warp == 0) {
    // ===================== TMA producer =====================
    int bs = 0, as = 0;
    uint32_t bph = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const Conv2Tile t = conv2_decode_tile<MT>(P, tile);
      int dxi = 0, cb = 0;
      for (int g = 0; g < ngroups; ++g) {
        for (int d = 0; d < G; ++d) {
          mbar_wait(&empty_bar[bs], bph ^ 1);
          if (elect_one()) {
            mbar_arrive_expect_tx(&full_bar[bs], B_TILE_BYTES + (d == 0 ? a_slot : 0));
            if (d == 0) {
              uint8_t* sa = smem_a + as * a_slot;
#pragma unroll
              for (int m = 0; m < MT; ++m)
                tma_load_4d(sa + m * a_box, &tmA, &full_bar[bs], cb * CONV_BK, t.x0 + 8 * m + (k3 ? dxi - 1 : 0),
                            t.y0 - (k3 ? 1 : 0), t.n);
            }
            const int tap = k3 ? d * 3 + dxi : 0;  // (ky, kx) = (d, dxi)
            tma_load_2d(smem_b + bs * B_SLOT, &tmB, &full_bar[bs], (tap * P.cin_blocks + cb) * CONV_BK, t.n_blk * BN);
          }
          __syncwarp();
          if (++bs == NB) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (++as == NA) as = 0;
        if (++cb == P.cin_blocks) {
          cb = 0;
          ++dxi;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    int bs = 0, as = 0;
    uint32_t bph = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * MT * BN;
      for (int g = 0; g < ngroups; ++g) {
        for (int d = 0; d < G; ++d) {
          mbar_wait(&full_bar[bs], bph);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t sa = smem_u32(smem_a + as * a_slot) + (k3 ? d * 1024 : 0);  // the dy view: one image row = 1024 B
            const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem_b + bs * B_SLOT));
#pragma unroll
            for (int m = 0; m < MT; ++m) {
              const uint64_t adesc = make_kmajor_sw128_desc(sa + m * a_box);
#pragma unroll
              for (int k = 0; k < CONV_BK / 8; ++k)
                umma_ss_tf32(d_tmem + m * BN, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), IDESC, (g | d | k) != 0);
            }
            umma_commit(&empty_bar[bs]);
            if (g == ngroups - 1 && d == G - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++bs == NB) {
            bs = 0;
            bph ^= 1;
          }
        }
        if (++as == NA) as = 0;
      }
      if (++acc == ACC) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue warps: warp (q, half) owns accumulator rows q*32..+32 and half of the 32-column steps ====
    constexpr int STEPS = BN / 32, STEPS0 = (STEPS + 1) / 2;
    const int q = warp & 3;
    const int ew = warp - 2;
    const int half = ew >> 2;
    const int step_lo = half ? STEPS0 : 0, step_hi = half ? STEPS : STEPS0;
    const uint32_t stg = smem_u32(smem_epi + ew * 4096);
    const float* __restrict__ resp = P.res;
    float* __restrict__ outp = P.out;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const Conv2Tile t = conv2_decode_tile<MT>(P, tile);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        // the 8 rows this lane stores: accumulator row r = q*32 + it*4 + lane/8 is pixel (x0 + 8m + r % 8, y0 + r / 8)
        int pix[8];  // N * H * W < 2^31 (checked by the launcher)
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rl = q * 32 + it * 4 + (lane >> 3);
          const int x = t.x0 + 8 * m + (rl & 7), y = t.y0 + (rl >> 3);
          pix[it] = (x < P.W && y < P.H) ? (t.n * P.H + y) * P.W + x : -1;
        }
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16) + (acc * MT + m) * BN;
#pragma unroll 1
        for (int step = step_lo; step < step_hi; ++step) {
          const int col = step * 32;
          const int n0 = t.n_blk * BN + col;
          if (n0 >= P.Cout) break;  // warp-uniform
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_row + col, r);
          const int c16 = lane & 7;
          const int gn = n0 + c16 * 4;
          const bool cok = gn < P.Cout;
          // residual rows are fetched while the TMEM read is in flight, all eight before any store (independent loads)
          float4 rv[8];
          if (resp) {
#pragma unroll
            for (int it = 0; it < 8; ++it)
              rv[it] = (cok && pix[it] >= 0) ? __ldg(reinterpret_cast<const float4*>(resp + (size_t)pix[it] * P.ldr + gn))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (P.bias && cok) bv = __ldg(reinterpret_cast<const float4*>(P.bias + gn));
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            sts_v4(stg + lane * 128 + ((j ^ (lane & 7)) << 4), make_uint4(r[j * 4], r[j * 4 + 1], r[j * 4 + 2], r[j * 4 + 3]));
          __syncwarp();
          if (cok) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const int rr = it * 4 + (lane >> 3);
              const uint4 v = lds_v4(stg + rr * 128 + ((c16 ^ (rr & 7)) << 4));
              if (pix[it] >= 0) {
                float4 o = make_float4(__uint_as_float(v.x) + bv.x, __uint_as_float(v.y) + bv.y, __uint_as_float(v.z) + bv.z,
                                       __uint_as_float(v.w) + bv.w);
                if (resp) {
                  o.x += rv[it].x; o.y += rv[it].y; o.z += rv[it].z; o.w += rv[it].w;
                }
                *reinterpret_cast<float4*>(outp + (size_t)pix[it] * P.ldo + gn) = o;
              }
            }
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == ACC) {
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

// ---------------------------------------------------------------------------------------------------------------------
// bandwidth-bound row kernels (fp32)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// QwenImageRMS_norm (autoencoder_kl_qwenimage.py:102-109): F.normalize(x, dim=channel) * sqrt(C) * gamma, then optional SiLU
// (:246-247,262-263,640-641).  A warp normalises 12 / CPL pixels per iteration (CPL = C / 32 values per lane and pixel):
// always 12 independent 128-byte-per-warp loads in flight per thread (one pixel at a time ran at 40 % of HBM bandwidth).
template <int CPL>
__global__ void __launch_bounds__(256) vae_rms_act_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                          float* __restrict__ y, long long rows, int silu) {
  constexpr int C = CPL * 32, PIX = 12 / CPL;
  const int lane = threadIdx.x & 31;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  float g[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) g[i] = __ldg(gamma + lane + i * 32);
  const float sc = sqrtf((float)C);
  for (long long r0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * PIX; r0 < rows; r0 += nwarps * PIX) {
    float v[PIX][CPL];
#pragma unroll
    for (int p = 0; p < PIX; ++p)
#pragma unroll
      for (int i = 0; i < CPL; ++i) v[p][i] = (r0 + p < rows) ? x[(r0 + p) * C + lane + i * 32] : 0.f;
#pragma unroll
    for (int p = 0; p < PIX; ++p) {
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < CPL; ++i) ss = fmaf(v[p][i], v[p][i], ss);
      ss = warp_sum(ss);
      const float s = sc / fmaxf(sqrtf(ss), 1e-12f);
      if (r0 + p < rows) {
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
          const float o = v[p][i] * s * g[i];
          y[(r0 + p) * C + lane + i * 32] = silu ? silu_exact(o) : o;
        }
      }
    }
  }
}

// nearest-exact x2 (QwenImageUpsample, :147-156): out[n, y, x, :] = in[n, y / 2, x / 2, :]; one float4 per thread
__global__ void vae_upsample2x_kernel(const float4* __restrict__ in, float4* __restrict__ out, int N, int H, int W, int C4) {
  const long long total = (long long)N * (2 * H) * (2 * W) * C4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % C4);
    long long p = i / C4;
    const int x = (int)(p % (2 * W));
    p /= 2 * W;
    const int y = (int)(p % (2 * H));
    const int n = (int)(p / (2 * H));
    out[i] = in[(((long long)n * H + (y >> 1)) * W + (x >> 1)) * C4 + c];
  }
}

// post_quant_conv (1x1x1, z_dim -> z_dim, :848) fused with NCHW -> NHWC; output padded to 32 channels (one K block of the
// conv_in implicit GEMM), channels >= Z written as zero.  One thread per pixel.
template <int Z>
__global__ void vae_post_quant_kernel(const float* __restrict__ z, const float* __restrict__ w, const float* __restrict__ b,
                                      float* __restrict__ out, int N, int HW) {
  __shared__ float sw[Z * Z + Z];
  for (int i = threadIdx.x; i < Z * Z + Z; i += blockDim.x) sw[i] = i < Z * Z ? w[i] : b[i - Z * Z];
  __syncthreads();
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)N * HW) return;
  const int n = (int)(p / HW), i = (int)(p - (long long)n * HW);
  float v[Z];
#pragma unroll
  for (int c = 0; c < Z; ++c) v[c] = z[((long long)n * Z + c) * HW + i];
  float4* o = reinterpret_cast<float4*>(out + p * 32);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (c4 * 4 < Z) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int co = c4 * 4 + e;
        float a = sw[Z * Z + co];
#pragma unroll
        for (int ci = 0; ci < Z; ++ci) a = fmaf(sw[co * Z + ci], v[ci], a);
        r[e] = a;
      }
    }
    o[c4] = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// encoder input: NCHW image [N, C <= 32, H, W] -> NHWC [N, H, W, 32] with channels >= C zero (one K block of conv_in)
__global__ void vae_image_to_nhwc_kernel(const float* __restrict__ img, float* __restrict__ out, int N, int C, int HW) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= (long long)N * HW) return;
  const int n = (int)(p / HW), i = (int)(p - (long long)n * HW);
  float v[32];
#pragma unroll
  for (int c = 0; c < 32; ++c) v[c] = c < C ? img[((long long)n * C + c) * HW + i] : 0.f;
  float4* o = reinterpret_cast<float4*>(out + p * 32);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) o[c4] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
}

// conv_out (:656; 3x3, C -> 3 channels) + clamp(-1, 1) (:857) + NHWC -> NCHW (and / or the uint8 post-process).  x is the
// normalised, SiLU-activated input.  FMA pipe, exact fp32 (three output channels are no tensor-core shape).  A block of 128
// threads owns a 32 x 16 pixel tile; per 32-channel chunk the (32 + 2) x (16 + 2) input patch is staged in shared memory
// (pixel stride 36 floats: conflict-free 128-bit reads across a quarter warp) with that chunk's weights [3][9][32]; a thread
// computes 4 vertically adjacent pixels x 3 channels (12 accumulators), holding the 9 weight vectors of one kernel column in
// registers while it walks the 6 input rows: 9.6 FMA per shared-memory load.
constexpr int CO_TX = 32, CO_TY = 16, CO_PITCH = 36;
constexpr int CO_PATCH = (CO_TX + 2) * (CO_TY + 2);
constexpr int CO_SMEM_BYTES = CO_PATCH * CO_PITCH * 4 + 3 * 9 * 32 * 4;

template <int C>
__global__ void __launch_bounds__(128) vae_conv_out_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ b, float* __restrict__ out,
                                                           uint8_t* __restrict__ out_u8, int N, int H, int W) {
  extern __shared__ float4 co_smem[];
  float* sx = reinterpret_cast<float*>(co_smem);             // [CO_PATCH][CO_PITCH]
  float* sw = sx + CO_PATCH * CO_PITCH;                      // [3][9][32]
  const int tid = threadIdx.x, tx = tid & 31, ty4 = tid >> 5;
  const int x0 = blockIdx.x * CO_TX, y0 = blockIdx.y * CO_TY, n = blockIdx.z;
  float acc[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    acc[j][0] = b[0];
    acc[j][1] = b[1];
    acc[j][2] = b[2];
  }
#pragma unroll 1
  for (int c0 = 0; c0 < C; c0 += 32) {
    __syncthreads();  // previous chunk fully consumed
    for (int i = tid; i < CO_PATCH * 8; i += 128) {
      const int p = i >> 3, c4 = i & 7;
      const int py = p / (CO_TX + 2), px = p - py * (CO_TX + 2);
      const int gy = y0 + py - 1, gx = x0 + px - 1;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = __ldg(reinterpret_cast<const float4*>(x + (((long long)n * H + gy) * W + gx) * C + c0) + c4);
      *reinterpret_cast<float4*>(sx + p * CO_PITCH + c4 * 4) = v;
    }
    for (int i = tid; i < 3 * 9 * 32; i += 128) {
      const int co_tap = i >> 5, c = i & 31;
      sw[i] = w[co_tap * C + c0 + c];
    }
    __syncthreads();
#pragma unroll 1
    for (int c4 = 0; c4 < 8; ++c4) {
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        float4 wv[3][3];  // [ky][co]
#pragma unroll
        for (int ky = 0; ky < 3; ++ky)
#pragma unroll
          for (int co = 0; co < 3; ++co)
            wv[ky][co] = *reinterpret_cast<const float4*>(sw + ((co * 9 + ky * 3 + kx) << 5) + c4 * 4);
#pragma unroll
        for (int r = 0; r < 6; ++r) {
          const float4 v = *reinterpret_cast<const float4*>(sx + ((ty4 * 4 + r) * (CO_TX + 2) + tx + kx) * CO_PITCH + c4 * 4);
#pragma unroll
          for (int ky = 0; ky < 3; ++ky) {
            const int j = r - ky;
            if (j >= 0 && j < 4) {
#pragma unroll
              for (int co = 0; co < 3; ++co) {
                const float4 u = wv[ky][co];
                acc[j][co] = fmaf(v.x, u.x, fmaf(v.y, u.y, fmaf(v.z, u.z, fmaf(v.w, u.w, acc[j][co]))));
              }
            }
          }
        }
      }
    }
  }
  const int xx = x0 + tx;
  if (xx >= W) return;
  const long long hw = (long long)H * W;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int yy = y0 + ty4 * 4 + j;
    if (yy >= H) break;
    const float a0 = fminf(fmaxf(acc[j][0], -1.f), 1.f);
    const float a1 = fminf(fmaxf(acc[j][1], -1.f), 1.f);
    const float a2 = fminf(fmaxf(acc[j][2], -1.f), 1.f);
    const long long pix = (long long)yy * W + xx;
    if (out) {
      const long long o = (long long)n * 3 * hw + pix;
      out[o] = a0;
      out[o + hw] = a1;
      out[o + 2 * hw] = a2;
    }
    if (out_u8) {
      // VaeImageProcessor.postprocess (the reference's post_process_func, pipeline_qwen_image.py:40-60): denormalise
      // (x / 2 + 0.5).clamp(0, 1), NHWC, (x * 255).round() -> uint8; the same fp32 operations, rint = numpy's half-to-even
      uint8_t* o = out_u8 + ((long long)n * hw + pix) * 3;
      o[0] = (uint8_t)rintf(__fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(a0, 0.5f), 0.5f), 0.f), 1.f), 255.f));
      o[1] = (uint8_t)rintf(__fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(a1, 0.5f), 0.5f), 0.f), 1.f), 255.f));
      o[2] = (uint8_t)rintf(__fmul_rn(fminf(fmaxf(__fadd_rn(__fmul_rn(a2, 0.5f), 0.5f), 0.f), 1.f), 255.f));
    }
  }
}

// in-place row softmax of fp32 scores (F.scaled_dot_product_attention of the mid-block attention, :321): one CTA per row.
// SMEM = 1: the row is staged in shared memory (one global read, one write); SMEM = 0: three passes over global memory
// (rows longer than the shared memory holds).  T = float4 when the row length and stride allow it.
__device__ __forceinline__ float vmax(float v) { return v; }
__device__ __forceinline__ float vmax(float4 v) { return fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)); }
__device__ __forceinline__ float vexp(float& v, float m, float sc) { v = expf((v - m) * sc); return v; }
__device__ __forceinline__ float vexp(float4& v, float m, float sc) {
  v.x = expf((v.x - m) * sc); v.y = expf((v.y - m) * sc); v.z = expf((v.z - m) * sc); v.w = expf((v.w - m) * sc);
  return (v.x + v.y) + (v.z + v.w);
}
__device__ __forceinline__ float vscale(float v, float a) { return v * a; }
__device__ __forceinline__ float4 vscale(float4 v, float a) { return make_float4(v.x * a, v.y * a, v.z * a, v.w * a); }

template <int SMEM, typename T>
__global__ void vae_softmax_rows_kernel(float* __restrict__ s, int cols, long long ld, float scale) {
  extern __shared__ float4 srow4[];
  T* srow = reinterpret_cast<T*>(srow4);
  T* row = reinterpret_cast<T*>(s + (long long)blockIdx.x * ld);
  const int n = cols / (int)(sizeof(T) / 4);
  __shared__ float red[32];
  __shared__ float bc;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {
    const T v = row[c];
    if (SMEM) srow[c] = v;
    m = fmaxf(m, vmax(v));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float v = lane < nw ? red[lane] : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0) bc = v;
  }
  __syncthreads();
  m = bc;
  float sum = 0.f;
  for (int c = threadIdx.x; c < n; c += blockDim.x) {  // every thread revisits the elements it staged itself
    T v = SMEM ? srow[c] : row[c];
    sum += vexp(v, m, scale);
    if (SMEM) srow[c] = v; else row[c] = v;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float v = lane < nw ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) bc = v;
  }
  __syncthreads();
  const float inv = 1.0f / bc;
  for (int c = threadIdx.x; c < n; c += blockDim.x) row[c] = vscale(SMEM ? srow[c] : row[c], inv);
}

// out[c, r] = in[r * ld_in + c]  (V^T for the P*V GEMM: both tcgen05 operands K-major)
__global__ void vae_transpose_kernel(const float* __restrict__ in, long long ld_in, float* __restrict__ out, int rows, int cols) {
  __shared__ float t[32][33];
  const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    t[i][threadIdx.x] = (r < rows && c < cols) ? in[(long long)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(long long)c * rows + r] = t[threadIdx.x][i];
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn vae_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int encode_f32(CUtensorMap* tm, int rank, const void* ptr, const cuuint64_t* gdim, const cuuint64_t* gstride_bytes,
                      const cuuint32_t* box, int pixel_stride = 1) {
  EncodeTiledFn enc = vae_encode_fn();
  if (!enc) return fail("cuTensorMapEncodeTiled unavailable (no CUDA driver / no GPU)");
  // element strides: a box dimension of n * s elements walked with stride s loads n elements (dims 1, 2 = x, y of a 4-D map)
  cuuint32_t estr[4] = {1, (cuuint32_t)pixel_stride, (cuuint32_t)pixel_stride, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(ptr), gdim, gstride_bytes, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "CUresult %d (rank %d, dims %llu %llu %llu %llu)", (int)r, rank, (unsigned long long)gdim[0],
             (unsigned long long)gdim[1], rank > 2 ? (unsigned long long)gdim[2] : 0ull, rank > 3 ? (unsigned long long)gdim[3] : 0ull);
    return fail("cuTensorMapEncodeTiled(fp32)", buf);
  }
  return 0;
}

template <typename K>
static int conv_smem_attr(K kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return fail("no CUDA device");
  if (!done[dev]) {
    QIMG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done[dev] = true;
  }
  return 0;
}

// 0 = v2 (shared vertical taps + 256-pixel tiles; default), 1 = v1 (one box per tap, 128-pixel tiles): env QIMG_VAE_CONV
static int g_conv_variant = -1;
static int conv_variant() {
  if (g_conv_variant < 0) {
    const char* e = getenv("QIMG_VAE_CONV");
    g_conv_variant = (e && atoi(e) == 1) ? 1 : 0;
  }
  return g_conv_variant;
}

}  // namespace qimg

using namespace qimg;

extern "C" {

int qimg_conv2d_nhwc_tf32(const float* x, int ldx, const float* w, int ldw, const float* bias, const float* res, int ldr,
                          float* out, int ldo, int N, int H, int W, int Cin, int Cout, int taps, qimg_stream_t stream) {
  if (!x || !w || !out) return fail("qimg_conv2d_nhwc_tf32: null pointer");
  if (taps != 1 && taps != 9) return fail("qimg_conv2d_nhwc_tf32: taps must be 1 (1x1) or 9 (3x3)");
  if (N < 1 || H < CONV_PY || W < CONV_PX) return fail("qimg_conv2d_nhwc_tf32: needs H >= 8 and W >= 16 (one 16 x 8 pixel patch)");
  if (Cin < 32 || (Cin & 3) || (taps == 9 && (Cin & 31)))
    return fail("qimg_conv2d_nhwc_tf32: Cin must be >= 32, a multiple of 4 (1x1) / of 32 (3x3: pad the channels)");
  if (Cout < 1 || (Cout & 3)) return fail("qimg_conv2d_nhwc_tf32: Cout must be a multiple of 4");
  const int cin_blocks = (Cin + CONV_BK - 1) / CONV_BK;
  const long long kcols = taps == 9 ? 9ll * cin_blocks * CONV_BK : Cin;
  if (ldx < Cin || (ldx & 3) || ldw < kcols || (ldw & 3) || ldo < Cout || (ldo & 3) || (res && (ldr < Cout || (ldr & 3))))
    return fail("qimg_conv2d_nhwc_tf32: leading dimensions must cover the row and be multiples of 4 floats");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(res) | reinterpret_cast<uintptr_t>(bias)) & 15)
    return fail("qimg_conv2d_nhwc_tf32: pointers must be 16-byte aligned");
  const int sms = device_sm_count();
  if (sms <= 0) return fail("no CUDA device");
  const int variant = conv_variant();
  // v2 tile: BN output channels (96 / 192 for the decoder's 96-, 192- and 384-wide layers: no padded MMA columns; 128 for
  // everything else), MT 8 x 16 pixel blocks per CTA (2 = a 16 x 16 patch sharing every weight tile, while that still gives
  // two waves of tiles)
  const int BN = variant == 1 ? CONV_BN : (Cout == 96 ? 96 : ((Cout == 192 || Cout == 384) ? 192 : CONV_BN));
  const int n_tiles = (Cout + BN - 1) / BN;
  const long long m2 = (long long)N * ((W + 15) / 16) * ((H + 15) / 16);
  const int MT = (variant == 0 && m2 * n_tiles >= 2ll * sms) ? 2 : 1;
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W * ldx * 4, (cuuint64_t)H * W * ldx * 4};
    cuuint32_t box1[4] = {CONV_BK, CONV_PX, CONV_PY, 1};
    cuuint32_t box2[4] = {CONV_BK, 8, (cuuint32_t)(taps == 9 ? 18 : 16), 1};
    if (encode_f32(&tmA, 4, x, gdim, gstr, variant == 1 ? box1 : box2)) return 1;
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)kcols, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)ldw * 4};
    cuuint32_t box[2] = {CONV_BK, (cuuint32_t)BN};
    if (encode_f32(&tmB, 2, w, gdim, gstr, box)) return 1;
  }
  ConvParams P;
  memset(&P, 0, sizeof P);
  P.N = N; P.H = H; P.W = W;
  P.cin_blocks = cin_blocks;
  P.taps = taps;
  P.stride = 1;
  P.Cout = Cout;
  P.ldo = ldo; P.ldr = ldr;
  if (variant == 1) {
    P.tiles_x = (W + CONV_PX - 1) / CONV_PX;
    P.tiles_y = (H + CONV_PY - 1) / CONV_PY;
  } else {
    P.tiles_x = (W + 8 * MT - 1) / (8 * MT);
    P.tiles_y = (H + 15) / 16;
  }
  P.n_tiles = n_tiles;
  const long long total = (long long)N * P.tiles_x * P.tiles_y * P.n_tiles;
  if (total > 0x7fffffffll || (long long)N * H * W > 0x7fffffffll) return fail("qimg_conv2d_nhwc_tf32: too many tiles / pixels");
  P.total_tiles = (int)total;
  P.bias = bias; P.res = res; P.out = out;
  const int grid = P.total_tiles < sms ? P.total_tiles : sms;
  cudaStream_t st = (cudaStream_t)stream;
#define QIMG_CONV2_LAUNCH(MT_, BN_, ACC_)                                                                  \
  do {                                                                                                     \
    static bool done[64] = {};                                                                             \
    if (conv_smem_attr(conv2_tf32_kernel<MT_, BN_, ACC_>, c2_smem_bytes(BN_), done)) return 1;             \
    conv2_tf32_kernel<MT_, BN_, ACC_><<<grid, CONV_THREADS, c2_smem_bytes(BN_), st>>>(tmA, tmB, P);        \
  } while (0)
  if (variant == 1) {
    static bool done[64] = {};
    if (conv_smem_attr(conv_tf32_kernel, CONV_SMEM_BYTES, done)) return 1;
    conv_tf32_kernel<<<grid, CONV_THREADS, CONV_SMEM_BYTES, st>>>(tmA, tmB, P);
  } else if (BN == 96) {
    if (MT == 2) QIMG_CONV2_LAUNCH(2, 96, 2); else QIMG_CONV2_LAUNCH(1, 96, 2);
  } else if (BN == 192) {
    if (MT == 2) QIMG_CONV2_LAUNCH(2, 192, 1); else QIMG_CONV2_LAUNCH(1, 192, 2);  // 2 x 192 x 2 columns exceed TMEM: one stage
  } else {
    if (MT == 2) QIMG_CONV2_LAUNCH(2, 128, 2); else QIMG_CONV2_LAUNCH(1, 128, 2);
  }
#undef QIMG_CONV2_LAUNCH
  QIMG_LAUNCH_CHECK("conv_tf32_kernel");
  return 0;
}

int qimg_conv2d_down2_nhwc_tf32(const float* x, int ldx, const float* w, int ldw, const float* bias, float* out, int ldo, int N,
                                int H_in, int W_in, int Cin, int Cout, qimg_stream_t stream) {
  if (!x || !w || !out) return fail("qimg_conv2d_down2_nhwc_tf32: null pointer");
  if (N < 1 || (H_in & 1) || (W_in & 1) || H_in < 2 * CONV_PY || W_in < 2 * CONV_PX)
    return fail("qimg_conv2d_down2_nhwc_tf32: needs even H_in >= 16 and W_in >= 32");
  if (Cin < 32 || (Cin & 31)) return fail("qimg_conv2d_down2_nhwc_tf32: Cin must be a multiple of 32");
  if (Cout < 1 || (Cout & 3)) return fail("qimg_conv2d_down2_nhwc_tf32: Cout must be a multiple of 4");
  const int H = H_in / 2, W = W_in / 2, cin_blocks = Cin / CONV_BK;
  const long long kcols = 9ll * Cin;
  if (ldx < Cin || (ldx & 3) || ldw < kcols || (ldw & 3) || ldo < Cout || (ldo & 3))
    return fail("qimg_conv2d_down2_nhwc_tf32: leading dimensions must cover the row and be multiples of 4 floats");
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out) |
       reinterpret_cast<uintptr_t>(bias)) & 15)
    return fail("qimg_conv2d_down2_nhwc_tf32: pointers must be 16-byte aligned");
  const int sms = device_sm_count();
  if (sms <= 0) return fail("no CUDA device");
  CUtensorMap tmA, tmB;
  {
    cuuint64_t gdim[4] = {(cuuint64_t)Cin, (cuuint64_t)W_in, (cuuint64_t)H_in, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)ldx * 4, (cuuint64_t)W_in * ldx * 4, (cuuint64_t)H_in * W_in * ldx * 4};
    cuuint32_t box[4] = {CONV_BK, 2 * CONV_PX, 2 * CONV_PY, 1};  // 16 x 8 pixels at element stride 2
    if (encode_f32(&tmA, 4, x, gdim, gstr, box, 2)) return 1;
  }
  {
    cuuint64_t gdim[2] = {(cuuint64_t)kcols, (cuuint64_t)Cout};
    cuuint64_t gstr[1] = {(cuuint64_t)ldw * 4};
    cuuint32_t box[2] = {CONV_BK, CONV_BN};
    if (encode_f32(&tmB, 2, w, gdim, gstr, box)) return 1;
  }
  ConvParams P;
  memset(&P, 0, sizeof P);
  P.N = N; P.H = H; P.W = W;
  P.cin_blocks = cin_blocks;
  P.taps = 9;
  P.stride = 2;
  P.Cout = Cout;
  P.ldo = ldo;
  P.tiles_x = (W + CONV_PX - 1) / CONV_PX;
  P.tiles_y = (H + CONV_PY - 1) / CONV_PY;
  P.n_tiles = (Cout + CONV_BN - 1) / CONV_BN;
  const long long total = (long long)N * P.tiles_x * P.tiles_y * P.n_tiles;
  if (total > 0x7fffffffll) return fail("qimg_conv2d_down2_nhwc_tf32: too many tiles");
  P.total_tiles = (int)total;
  P.bias = bias; P.out = out;
  static bool done[64] = {};
  if (conv_smem_attr(conv_tf32_kernel, CONV_SMEM_BYTES, done)) return 1;
  const int grid = P.total_tiles < sms ? P.total_tiles : sms;
  conv_tf32_kernel<<<grid, CONV_THREADS, CONV_SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmB, P);
  QIMG_LAUNCH_CHECK("conv_tf32_kernel(stride 2)");
  return 0;
}

int qimg_vae_image_to_nhwc(const float* img, float* out, int N, int C, int H, int W, qimg_stream_t stream) {
  if (!img || !out || N < 1 || C < 1 || C > 32 || H < 1 || W < 1) return fail("qimg_vae_image_to_nhwc: bad arguments (C <= 32)");
  const long long px = (long long)N * H * W;
  vae_image_to_nhwc_kernel<<<(unsigned)((px + 255) / 256), 256, 0, (cudaStream_t)stream>>>(img, out, N, C, H * W);
  QIMG_LAUNCH_CHECK("vae_image_to_nhwc_kernel");
  return 0;
}

int qimg_set_vae_conv_variant(int variant) {
  if (variant < 0 || variant > 1) return fail("qimg_set_vae_conv_variant: 0 (shared vertical taps, 256-pixel tiles) or 1 (one box per tap)");
  g_conv_variant = variant;
  return 0;
}

int qimg_vae_rms_act(const float* x, const float* gamma, float* y, long long rows, int C, int silu, qimg_stream_t stream) {
  if (!x || !gamma || !y || rows < 1) return fail("qimg_vae_rms_act: bad arguments");
  if (C != 96 && C != 192 && C != 384) return fail("qimg_vae_rms_act: C must be 96, 192 or 384 (the decoder's channel widths)");
  const int sms = device_sm_count();
  if (sms <= 0) return fail("no CUDA device");
  const int pix = 12 / (C / 32);
  long long blocks = (rows + 8ll * pix - 1) / (8ll * pix);
  if (blocks > 8ll * sms) blocks = 8ll * sms;
  cudaStream_t st = (cudaStream_t)stream;
  if (C == 96) vae_rms_act_kernel<3><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, y, rows, silu);
  else if (C == 192) vae_rms_act_kernel<6><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, y, rows, silu);
  else vae_rms_act_kernel<12><<<(unsigned)blocks, 256, 0, st>>>(x, gamma, y, rows, silu);
  QIMG_LAUNCH_CHECK("vae_rms_act_kernel");
  return 0;
}

int qimg_vae_upsample2x(const float* x, float* out, int N, int H, int W, int C, qimg_stream_t stream) {
  if (!x || !out || N < 1 || H < 1 || W < 1 || C < 4 || (C & 3)) return fail("qimg_vae_upsample2x: bad arguments (C % 4 == 0)");
  const long long total = (long long)N * 4 * H * W * (C / 4);
  long long blocks = (total + 255) / 256;
  const int sms = device_sm_count();
  if (sms > 0 && blocks > (long long)sms * 32) blocks = (long long)sms * 32;
  vae_upsample2x_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x),
                                                                             reinterpret_cast<float4*>(out), N, H, W, C / 4);
  QIMG_LAUNCH_CHECK("vae_upsample2x_kernel");
  return 0;
}

int qimg_vae_post_quant(const float* z, const float* w, const float* b, float* out, int N, int H, int W, int z_dim,
                        qimg_stream_t stream) {
  if (!z || !w || !b || !out || N < 1 || H < 1 || W < 1) return fail("qimg_vae_post_quant: bad arguments");
  if (z_dim != 16) return fail("qimg_vae_post_quant: z_dim must be 16 (AutoencoderKLQwenImage)");
  const long long px = (long long)N * H * W;
  vae_post_quant_kernel<16><<<(unsigned)((px + 127) / 128), 128, 0, (cudaStream_t)stream>>>(z, w, b, out, N, H * W);
  QIMG_LAUNCH_CHECK("vae_post_quant_kernel");
  return 0;
}

int qimg_vae_conv_out(const float* x, const float* w, const float* b, float* out, uint8_t* out_u8, int N, int H, int W, int C,
                      qimg_stream_t stream) {
  if (!x || !w || !b || (!out && !out_u8) || N < 1 || H < 1 || W < 1) return fail("qimg_vae_conv_out: bad arguments");
  if (C != 96) return fail("qimg_vae_conv_out: C must be 96 (base_dim of AutoencoderKLQwenImage)");
  if (N > 65535) return fail("qimg_vae_conv_out: N too large");
  static bool done[64] = {};
  if (conv_smem_attr(vae_conv_out_kernel<96>, CO_SMEM_BYTES, done)) return 1;
  dim3 grid((W + CO_TX - 1) / CO_TX, (H + CO_TY - 1) / CO_TY, N);
  vae_conv_out_kernel<96><<<grid, 128, CO_SMEM_BYTES, (cudaStream_t)stream>>>(x, w, b, out, out_u8, N, H, W);
  QIMG_LAUNCH_CHECK("vae_conv_out_kernel");
  return 0;
}

int qimg_vae_softmax_rows(float* s, int rows, int cols, long long ld, float scale, qimg_stream_t stream) {
  if (!s || rows < 1 || cols < 1 || ld < cols) return fail("qimg_vae_softmax_rows: bad arguments");
  const size_t bytes = (size_t)cols * 4;
  const bool vec = !(cols & 3) && !(ld & 3) && !(reinterpret_cast<uintptr_t>(s) & 15);
  const bool in_smem = bytes <= 100 * 1024;  // two rows per SM
  cudaStream_t st = (cudaStream_t)stream;
  if (in_smem && vec) {
    static bool done[64] = {};
    if (conv_smem_attr(vae_softmax_rows_kernel<1, float4>, 100 * 1024, done)) return 1;
    vae_softmax_rows_kernel<1, float4><<<rows, 1024, bytes, st>>>(s, cols, ld, scale);
  } else if (in_smem) {
    static bool done[64] = {};
    if (conv_smem_attr(vae_softmax_rows_kernel<1, float>, 100 * 1024, done)) return 1;
    vae_softmax_rows_kernel<1, float><<<rows, 1024, bytes, st>>>(s, cols, ld, scale);
  } else if (vec) {
    vae_softmax_rows_kernel<0, float4><<<rows, 1024, 0, st>>>(s, cols, ld, scale);
  } else {
    vae_softmax_rows_kernel<0, float><<<rows, 1024, 0, st>>>(s, cols, ld, scale);
  }
  QIMG_LAUNCH_CHECK("vae_softmax_rows_kernel");
  return 0;
}

int qimg_vae_transpose(const float* in, long long ld_in, float* out, int rows, int cols, qimg_stream_t stream) {
  if (!in || !out || rows < 1 || cols < 1 || ld_in < cols) return fail("qimg_vae_transpose: bad arguments");
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  if (grid.y > 65535) return fail("qimg_vae_transpose: too many rows");
  vae_transpose_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ld_in, out, rows, cols);
  QIMG_LAUNCH_CHECK("vae_transpose_kernel");
  return 0;
}

}  // extern "C"
