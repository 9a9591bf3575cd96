// Tensor-parallel reduction over NVLink PEER MEMORY, fused with the bias + gate + residual epilogue
// (no NCCL on the data path).  Used by the engine's TP mode when qimg_engine_set_tp_p2p() was called.
//
// Each rank's row-parallel GEMM leaves bf16 partial sums in its own `part` buffer.  One kernel per rank then
//   * reads the partial sums of ITS slice of rows from every rank (P2P loads over NVLink, fp32 accumulation,
//     a single rounding to bf16 — tighter than a bf16 ring all-reduce),
//   * applies  x = x + gate * (sum + bias)  (qwen_image_transformer.py:586-587,592,597), and
//   * stores the updated rows of the residual stream x into EVERY rank's copy (P2P stores),
// i.e. reduce-scatter + epilogue + all-gather in one pass: 2 (P-1)/P * rows * D * 2 B cross NVLink per rank.
// Two cross-GPU barriers (flag arrays in peer memory, release/acquire at system scope) bracket it: all partial
// sums written before anyone reads them; all x rows written before anyone's next kernel reads x.
#include "../../include/qimg_b200.h"

#include <cstring>

#include "qimg_common.cuh"
#include "qimg_host.cuh"

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
    if (clock64() - t0 > 8000000000LL) {  // ~4 s: a peer died; record and bail out instead of hanging the GPU
      reinterpret_cast<int*>(local)[16] = 1;
      break;
    }
  }
}

template <int P>
__global__ void __launch_bounds__(256)
tp_reduce_gate_res_kernel(PeerPtrs part, PeerPtrs x, const bf16* __restrict__ bias, const bf16* __restrict__ gate,
                          long long vec_begin, long long vec_end, int D, int rows_per_batch, long long gate_stride, int rank) {
  const int dv = D >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = vec_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < vec_end; i += stride) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint4 v[P];
#pragma unroll
    for (int p = 0; p < P; ++p) v[p] = ldg_nc_v4(reinterpret_cast<const bf16*>(part.p[p]) + i * 8);  // P2P loads in flight together
#pragma unroll
    for (int p = 0; p < P; ++p) {
      const uint32_t w[4] = {v[p].x, v[p].y, v[p].z, v[p].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[2 * k] += bf16lo(w[k]);
        acc[2 * k + 1] += bf16hi(w[k]);
      }
    }
    const long long row = i / dv;
    const int col = (int)(i - row * dv) << 3;
    const long long b = row / rows_per_batch;
    const uint4 xv = ldg_v4(reinterpret_cast<const bf16*>(x.p[rank]) + i * 8);
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate + b * gate_stride + col));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + col));
    uint4 o;
    o.x = badd2(xv.x, bmul2(gv.x, badd2(pack_bf16x2(acc[0], acc[1]), bv.x)));
    o.y = badd2(xv.y, bmul2(gv.y, badd2(pack_bf16x2(acc[2], acc[3]), bv.y)));
    o.z = badd2(xv.z, bmul2(gv.z, badd2(pack_bf16x2(acc[4], acc[5]), bv.z)));
    o.w = badd2(xv.w, bmul2(gv.w, badd2(pack_bf16x2(acc[6], acc[7]), bv.w)));
#pragma unroll
    for (int p = 0; p < P; ++p) stg_v4(reinterpret_cast<bf16*>(x.p[p]) + i * 8, o);  // all-gather: every rank's x
  }
}

int tp_p2p_barrier(void* const* flags, int P, int rank, cudaStream_t st) {
  PeerPtrs f;
  memset(&f, 0, sizeof f);
  for (int p = 0; p < P; ++p) f.p[p] = flags[p];
  tp_barrier_kernel<<<1, 32, 0, st>>>(f, P, rank);
  QIMG_LAUNCH_CHECK("tp_barrier_kernel");
  return 0;
}

int tp_p2p_reduce(void* const* part, void* const* x, const void* bias, const void* gate, int rows, int D, int rows_per_batch,
                  long long gate_stride, int P, int rank, cudaStream_t st) {
  PeerPtrs pp, xx;
  memset(&pp, 0, sizeof pp);
  memset(&xx, 0, sizeof xx);
  for (int p = 0; p < P; ++p) {
    pp.p[p] = part[p];
    xx.p[p] = x[p];
  }
  // this rank's contiguous slice of the rows (balanced, first rows%P ranks get one more)
  const int base = rows / P, extra = rows % P;
  const long long r0 = (long long)rank * base + (rank < extra ? rank : extra);
  const long long r1 = r0 + base + (rank < extra ? 1 : 0);
  const long long v0 = r0 * (D / 8), v1 = r1 * (D / 8);
  if (v1 <= v0) return 0;
  long long blocks = (v1 - v0 + 255) / 256;
  const long long cap = (long long)device_sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const bf16* b = (const bf16*)bias;
  const bf16* g = (const bf16*)gate;
  switch (P) {
    case 2: tp_reduce_gate_res_kernel<2><<<(int)blocks, 256, 0, st>>>(pp, xx, b, g, v0, v1, D, rows_per_batch, gate_stride, rank); break;
    case 4: tp_reduce_gate_res_kernel<4><<<(int)blocks, 256, 0, st>>>(pp, xx, b, g, v0, v1, D, rows_per_batch, gate_stride, rank); break;
    case 8: tp_reduce_gate_res_kernel<8><<<(int)blocks, 256, 0, st>>>(pp, xx, b, g, v0, v1, D, rows_per_batch, gate_stride, rank); break;
    default: return fail("tp_p2p_reduce: tp_size must be 2, 4 or 8");
  }
  QIMG_LAUNCH_CHECK("tp_reduce_gate_res_kernel");
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
