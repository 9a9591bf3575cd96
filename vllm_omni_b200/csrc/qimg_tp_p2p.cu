// Tensor parallelism over NVLink PEER MEMORY (no NCCL on the data path).  Used by the engine's TP mode when
// qimg_engine_set_tp_p2p() was called.  Sequence-parallel epilogue design (round 2):
//
//   row-parallel GEMM (to_out / net.2, K sharded)      csrc/qimg_gemm.cuh, EPI_PARTIAL_F32
//       the tcgen05 epilogue pushes each fp32 accumulator tile over NVLink into the receive buffer of the rank that OWNS
//       the rows (reduce-scatter fused into the GEMM; overlapped with the next tile's main loop)
//   barrier                                             tp_barrier_kernel (flags in peer memory, release/acquire.sys)
//   tp_reduce_ln_push_kernel  (this file, one launch for image + text rows, own rows only = rows / P)
//       x = x + gate * (sum_p partial_p + bias)         fp32 sum of fp32 partials, ONE rounding — qwen_image_transformer.py:586-587,592,597
//       xm = LN(x) * (1 + scale) + shift                the NEXT AdaLayerNorm (layers/adalayernorm.py:94-102), row still in registers
//       xm row -> every rank's activation buffer        all-gather fused into the same kernel (P2P stores)
//   barrier
//   next column-parallel GEMM reads the full xm
// so the residual stream x lives row-sharded (each rank updates only its rows), the two AdaLN passes run on rows / P per
// rank, and per reduction a rank sends (P-1)/P * rows * D * (4 + 2) bytes.  (Round 1: bf16 partial sums read remotely AFTER
// the GEMM had finished, x all-gathered, both LayerNorms replicated on every rank: 1.37x at TP=2, 9.8e-3 from the
// single-GPU result at L=8.)
#include "../../include/qimg_b200.h"

#include <cstring>

#include "qimg_common.cuh"
#include "qimg_host.cuh"
#include "qimg_tp.h"

namespace qimg {

struct PeerPtrs {
  void* p[8];
};

__device__ __forceinline__ void st_release_sys(uint32_t* ptr, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* ptr) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}

// flags buffer of a rank (>= 128 B, in that rank's memory): u32 slots [0,8) = last epoch rank r has reached, slot 16 = int
// time-out flag, slot 17 = this rank's barrier counter.  The epoch lives in device memory (not a launch argument) so the
// sequence is CUDA-graph capturable and survives engine re-creation.  Thread p publishes this rank's epoch into rank
// p's array, then waits until rank p's arrival shows up locally.
__global__ void tp_barrier_kernel(PeerPtrs flags, int P, int rank) {
  uint32_t* local = reinterpret_cast<uint32_t*>(flags.p[rank]);
  uint32_t epoch = 0;
  if (threadIdx.x == 0) {
    epoch = local[17] + 1;
    local[17] = epoch;
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  const int p = threadIdx.x;
  if (p >= P) return;
  st_release_sys(reinterpret_cast<uint32_t*>(flags.p[p]) + rank, epoch);
  const uint32_t* mine = local + p;
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
    if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer died or the ranks issued different launch sequences
      reinterpret_cast<int*>(local)[16] = 1;  // sticky marker (qimg_engine_p2p_error) ...
      __threadfence_system();
      __trap();  // ... and a fatal error: continuing would reduce rows the peers have not finished writing
    }
  }
}

// ---- fused reduce + bias + gate + residual + next AdaLayerNorm + all-gather (round 2) ----------------------------------
struct ReduceSeg {             // one stream (image or text rows) of a reduction
  const float* recv;           // local receive buffer: [P][recv_rows][D] fp32, this stream's rows start at row_off
  int recv_rows, row_off;
  bf16* x;                     // local residual stream [M, D]; only rows [r0, r1) are read / written
  int r0, r1, rows_per_batch;
  const bf16* bias;            // [D] bias of the row-parallel linear
  const bf16* gate;            // gate[b * gate_stride + n]
  const bf16* shift;           // modulation of the NEXT LayerNorm: shift[b * mod_stride + n], scale likewise
  const bf16* scale;
  long long gate_stride, mod_stride;
  PeerPtrs xm;                 // every rank's modulated-activation buffer [M, D] of this stream
  int blocks;                  // thread blocks (4 rows each) assigned to this segment
};
struct ReduceParams {
  ReduceSeg seg[2];
  float eps;
};

template <int P, int NCH>
__global__ void __launch_bounds__(128) tp_reduce_ln_push_kernel(const __grid_constant__ ReduceParams prm) {
  constexpr int D = NCH * 256;
  const int si = (int)blockIdx.x < prm.seg[0].blocks ? 0 : 1;
  const ReduceSeg& sg = prm.seg[si];
  const int lane = threadIdx.x & 31;
  const int row = sg.r0 + ((int)blockIdx.x - (si ? prm.seg[0].blocks : 0)) * 4 + (threadIdx.x >> 5);
  if (row >= sg.r1) return;
  const int b = row / sg.rows_per_batch;
  const size_t roff = (size_t)row * D + lane * 8;
  const bf16* gt = sg.gate + (size_t)b * sg.gate_stride + lane * 8;
  const float* rv = sg.recv + ((size_t)sg.row_off + (row - sg.r0)) * D + lane * 8;
  const size_t src_stride = (size_t)sg.recv_rows * D;
  uint64_t c[NCH * 4];  // the updated residual row as fp32 pairs
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint4 lo[P], hi[P];
#pragma unroll
    for (int p = 0; p < P; ++p) {  // all sources' loads in flight together
      lo[p] = ldg_nc_v4(rv + p * src_stride + i * 256);
      hi[p] = ldg_nc_v4(rv + p * src_stride + i * 256 + 4);
    }
#pragma unroll
    for (int p = 0; p < P; ++p) {
      acc[0] += __uint_as_float(lo[p].x); acc[1] += __uint_as_float(lo[p].y);
      acc[2] += __uint_as_float(lo[p].z); acc[3] += __uint_as_float(lo[p].w);
      acc[4] += __uint_as_float(hi[p].x); acc[5] += __uint_as_float(hi[p].y);
      acc[6] += __uint_as_float(hi[p].z); acc[7] += __uint_as_float(hi[p].w);
    }
    const uint4 xv = ldg_v4(sg.x + roff + i * 256);
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gt + i * 256));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(sg.bias + lane * 8 + i * 256));
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      // y = bf16(sum + bias): the Linear's bf16 output, exactly the single-GPU epilogue; x = bf16(x + bf16(gate * y))
      const uint32_t y = pack_bf16x2(acc[2 * k] + bf16lo(bw[k]), acc[2 * k + 1] + bf16hi(bw[k]));
      o[k] = badd2(xw[k], bmul2(gw[k], y));
      c[i * 4 + k] = ew_pack2(o[k] << 16, o[k] & 0xffff0000u);
    }
    stg_v4(sg.x + roff + i * 256, make_uint4(o[0], o[1], o[2], o[3]));
  }
  // ---- the next AdaLayerNorm on the row held in registers (same arithmetic as ln_modulate_fast_kernel) ----
  uint64_t s0 = 0, s1 = 0;
#pragma unroll
  for (int i = 0; i < NCH * 4; i += 2) {
    s0 = ew_add2(s0, c[i]);
    s1 = ew_add2(s1, c[i + 1]);
  }
  const float mean = warp_sum(ew_hsum2(ew_add2(s0, s1))) * (1.0f / (float)D);
  const uint64_t nmean2 = ew_splat2(-mean);
  uint64_t q0 = 0, q1 = 0;
#pragma unroll
  for (int i = 0; i < NCH * 4; i += 2) {
    c[i] = ew_add2(c[i], nmean2);
    c[i + 1] = ew_add2(c[i + 1], nmean2);
    q0 = ew_fma2(c[i], c[i], q0);
    q1 = ew_fma2(c[i + 1], c[i + 1], q1);
  }
  const float rstd = rsqrtf(warp_sum(ew_hsum2(ew_add2(q0, q1))) * (1.0f / (float)D) + prm.eps);
  const uint64_t rstd2 = ew_splat2(rstd), zero2 = 0;
  const bf16* sh = sg.shift + (size_t)b * sg.mod_stride + lane * 8;
  const bf16* sc = sg.scale + (size_t)b * sg.mod_stride + lane * 8;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const uint4 shv = __ldg(reinterpret_cast<const uint4*>(sh + i * 256));
    const uint4 scv = __ldg(reinterpret_cast<const uint4*>(sc + i * 256));
    const uint32_t s1w[4] = {shv.x, shv.y, shv.z, shv.w}, s2w[4] = {scv.x, scv.y, scv.z, scv.w};
    uint32_t o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t lo, hi;
      ew_unpack2(ew_fma2(c[i * 4 + k], rstd2, zero2), lo, hi);
      const uint32_t n = pack_bf16x2(__uint_as_float(lo), __uint_as_float(hi));
      o[k] = badd2(bmul2(n, badd2(0x3F803F80u, s2w[k])), s1w[k]);
    }
    const uint4 ov = make_uint4(o[0], o[1], o[2], o[3]);
#pragma unroll
    for (int p = 0; p < P; ++p) stg_v4(reinterpret_cast<bf16*>(sg.xm.p[p]) + roff + i * 256, ov);  // all-gather over NVLink
  }
}

template <int P>
static int launch_reduce_ln(const ReduceParams& prm, int D, cudaStream_t st) {
  const int blocks = prm.seg[0].blocks + prm.seg[1].blocks;
  if (blocks <= 0) return 0;
  switch (D) {
    case 3072: tp_reduce_ln_push_kernel<P, 12><<<blocks, 128, 0, st>>>(prm); break;
    case 1024: tp_reduce_ln_push_kernel<P, 4><<<blocks, 128, 0, st>>>(prm); break;
    case 256: tp_reduce_ln_push_kernel<P, 1><<<blocks, 128, 0, st>>>(prm); break;
    default: return fail("tp_reduce_ln_push: hidden size must be 256, 1024 or 3072 (instantiated row lengths)");
  }
  QIMG_LAUNCH_CHECK("tp_reduce_ln_push_kernel");
  return 0;
}

// balanced contiguous split of `rows` over P owners (the same rule as the GEMM's EPI_PARTIAL_F32 epilogue)
static inline void own_range(int rows, int P, int rank, int* r0, int* r1) {
  const int base = rows / P, extra = rows % P;
  *r0 = rank * base + (rank < extra ? rank : extra);
  *r1 = *r0 + base + (rank < extra ? 1 : 0);
}

int tp_p2p_reduce_ln_push(const TpReduceArgs& a, cudaStream_t st) {
  ReduceParams prm;
  memset(&prm, 0, sizeof prm);
  prm.eps = a.eps;
  for (int s = 0; s < 2; ++s) {
    ReduceSeg& g = prm.seg[s];
    g.recv = (const float*)a.recv_local;
    g.recv_rows = a.recv_rows;
    g.row_off = a.row_off[s];
    g.x = (bf16*)a.x[s];
    own_range(a.rows[s], a.P, a.rank, &g.r0, &g.r1);
    g.rows_per_batch = a.rows_per_batch[s];
    g.bias = (const bf16*)a.bias[s];
    g.gate = (const bf16*)a.gate[s];
    g.shift = (const bf16*)a.shift[s];
    g.scale = (const bf16*)a.scale[s];
    g.gate_stride = a.gate_stride[s];
    g.mod_stride = a.mod_stride[s];
    for (int p = 0; p < a.P; ++p) g.xm.p[p] = a.xm[s][p];
    g.blocks = (g.r1 - g.r0 + 3) / 4;
  }
  switch (a.P) {
    case 2: return launch_reduce_ln<2>(prm, a.D, st);
    case 4: return launch_reduce_ln<4>(prm, a.D, st);
    case 8: return launch_reduce_ln<8>(prm, a.D, st);
  }
  return fail("tp_reduce_ln_push: tp_size must be 2, 4 or 8");
}

__global__ void __launch_bounds__(256) tp_push_rows_kernel(const bf16* __restrict__ src, PeerPtrs dst, long long n_vec, int P) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 v = ldg_nc_v4(src + i * 8);
    for (int p = 0; p < P; ++p) stg_v4(reinterpret_cast<bf16*>(dst.p[p]) + i * 8, v);
  }
}

int tp_p2p_push_rows(const void* src, void* const* dst, long long n, int P, cudaStream_t st) {
  if (n <= 0) return 0;
  if (n % 8) return fail("tp_p2p_push_rows: element count must be a multiple of 8");
  PeerPtrs d;
  memset(&d, 0, sizeof d);
  for (int p = 0; p < P; ++p) d.p[p] = dst[p];
  long long blocks = (n / 8 + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  tp_push_rows_kernel<<<(int)blocks, 256, 0, st>>>((const bf16*)src, d, n / 8, P);
  QIMG_LAUNCH_CHECK("tp_push_rows_kernel");
  return 0;
}

int tp_p2p_barrier(void* const* flags, int P, int rank, cudaStream_t st) {
  PeerPtrs f;
  memset(&f, 0, sizeof f);
  for (int p = 0; p < P; ++p) f.p[p] = flags[p];
  tp_barrier_kernel<<<1, 32, 0, st>>>(f, P, rank);
  QIMG_LAUNCH_CHECK("tp_barrier_kernel");
  return 0;
}

}  // namespace qimg

using namespace qimg;

extern "C" {

int qimg_p2p_alloc(size_t bytes, void** out) {
  if (!out) return fail("qimg_p2p_alloc: null");
  QIMG_CUDA_CHECK(cudaMalloc(out, bytes));
  QIMG_CUDA_CHECK(cudaMemset(*out, 0, bytes));
  return 0;
}
int qimg_p2p_free(void* ptr) {
  QIMG_CUDA_CHECK(cudaFree(ptr));
  return 0;
}
int qimg_ipc_get_handle(const void* dev_ptr, void* handle64) {
  cudaIpcMemHandle_t h;
  QIMG_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t size");
  memcpy(handle64, &h, 64);
  return 0;
}
int qimg_ipc_open_handle(const void* handle64, void** out) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  QIMG_CUDA_CHECK(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int qimg_ipc_close_handle(void* ptr) {
  QIMG_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return 0;
}

}  // extern "C"
