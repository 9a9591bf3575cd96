// C-ABI entry points (include/qimg_b200.h) and host launchers for the sm_100a kernels.
#include "../../include/qimg_b200.h"

#include <cstdlib>
#include <cstring>
#include <vector>

#include "qimg_elementwise.cuh"
#include "qimg_fmha.cuh"
#include "qimg_fmha4.cuh"
#include "qimg_fmha6.cuh"
#include "qimg_gemm.cuh"
#include "qimg_gemm2.cuh"
#include "qimg_host.cuh"

namespace qimg {

thread_local std::string g_last_error;
static thread_local const int* g_launch_predicate = nullptr;
void set_launch_predicate(const int* flag) { g_launch_predicate = flag; }
const int* launch_predicate() { return g_launch_predicate; }
std::atomic<long long> g_launch_count{0};

constexpr int kMaxDevices = 64;

static int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
  return dev;
}

int device_sm_count() {
  static int cached[kMaxDevices] = {};
  const int dev = current_device();
  if (dev < 0) return -1;
  if (cached[dev] > 0) return cached[dev];
  int n = 0;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  cached[dev] = n;
  return n;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): set it once for each device this process
// launches on (`done` is the launcher's own static table)
template <typename K>
static int ensure_smem_attr(K kernel, int smem_bytes, bool (&done)[kMaxDevices]) {
  const int dev = current_device();
  if (dev < 0) return fail("no CUDA device");
  if (!done[dev]) {
    QIMG_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    done[dev] = true;
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// per-launch CUDA-event profiling of the two tensor-core kernels (bench.py roofline numbers)
// ------------------------------------------------------------------------------------------
struct ProfRec {
  cudaEvent_t a, b;
  double flops;
};
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;                      // guards g_prof / g_event_pool (launches may come from several host threads)
static std::vector<ProfRec> g_prof[2];            // 0 = gemm, 1 = fmha
static std::vector<cudaEvent_t> g_event_pool;

static cudaEvent_t prof_event() {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
struct ProfScope {
  int kind;
  cudaStream_t st;
  ProfRec rec;
  bool on;
  ProfScope(int kind_, double flops, cudaStream_t st_) : kind(kind_), st(st_), on(g_prof_on.load()) {
    if (on) {
      rec.a = prof_event();
      rec.b = prof_event();
      rec.flops = flops;
      cudaEventRecord(rec.a, st);
    }
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(rec.b, st);
      std::lock_guard<std::mutex> lk(g_prof_mu);
      g_prof[kind].push_back(rec);
    }
  }
};

// ------------------------------------------------------------------------------------------
// TMA descriptor cache
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
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

struct TmapKey {
  const void* ptr;
  uint64_t cols, rows, batches;
  uint32_t box_rows;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && cols == o.cols && rows == o.rows && batches == o.batches && box_rows == o.box_rows;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.cols * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.rows * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= (k.batches * 1315423911ull + k.box_rows) + (h << 6) + (h >> 2);
    return h;
  }
};
static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> g_tmaps;
static std::mutex g_tmap_mu;

static const CUtensorMap* get_tmap(const void* ptr, uint64_t cols, uint64_t rows, uint64_t batches, uint32_t box_rows) {
  TmapKey key{ptr, cols, rows, batches, box_rows};
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  auto it = g_tmaps.find(key);
  if (it != g_tmaps.end()) return &it->second;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    fail("cuTensorMapEncodeTiled unavailable (no CUDA driver / no GPU)");
    return nullptr;
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (cols % 8)) {
    fail("TMA operand must be 16-byte aligned with a row length that is a multiple of 8 elements");
    return nullptr;
  }
  CUtensorMap tm;
  const bool is3d = batches > 0;
  cuuint64_t gdim[3] = {cols, rows, is3d ? batches : 1};
  cuuint64_t gstride[2] = {cols * 2, cols * rows * 2};
  cuuint32_t box[3] = {64, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, is3d ? 3 : 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[96];
    snprintf(buf, sizeof buf, "CUresult %d (cols=%llu rows=%llu box_rows=%u)", (int)r, (unsigned long long)cols,
             (unsigned long long)rows, box_rows);
    fail("cuTensorMapEncodeTiled", buf);
    return nullptr;
  }
  auto ins = g_tmaps.emplace(key, tm);
  return &ins.first->second;
}
// Bounded cache: descriptors are keyed by raw pointer and activation buffers come and go with the caller's allocator.
// Called at the START of a launcher (never between two get_tmap calls of one launch, whose returned pointers must stay
// valid until the kernel is enqueued — the descriptor is copied into the kernel's parameter space at launch).
void tmap_cache_trim() {
  std::lock_guard<std::mutex> lk(g_tmap_mu);
  if (g_tmaps.size() >= 4096) g_tmaps.clear();
}
const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows) {
  return get_tmap(ptr, cols, rows, 0, box_rows);
}
const CUtensorMap* get_tmap_3d(const void* ptr, uint64_t cols, uint64_t rows, uint64_t batches, uint32_t box_rows) {
  return get_tmap(ptr, cols, rows, batches, box_rows);
}

// ------------------------------------------------------------------------------------------
// GEMM launcher
// ------------------------------------------------------------------------------------------
template <int BN, int EPI>
static int launch_gemm_inst(const CUtensorMap* tA[2], const CUtensorMap* tB[2], const GemmParams& prm, cudaStream_t st) {
  static bool attr_done[kMaxDevices] = {};
  constexpr int smem = gemm_smem_bytes<BN>();
  if (ensure_smem_attr(gemm_umma_kernel<BN, EPI>, smem, attr_done)) return 1;
  int sms = device_sm_count();
  if (sms <= 0) return fail("no CUDA device");
  int grid = prm.total_tiles < sms ? prm.total_tiles : sms;
  gemm_umma_kernel<BN, EPI><<<grid, GEMM_THREADS, smem, st>>>(*tA[0], *tB[0], *tA[1], *tB[1], prm);
  QIMG_LAUNCH_CHECK("gemm_umma_kernel");
  return 0;
}

template <int EPI>
static int launch_gemm2_inst(const CUtensorMap* tA[2], const CUtensorMap* tB[2], const GemmParams& prm, cudaStream_t st) {
  static bool attr_done[kMaxDevices] = {};
  if (ensure_smem_attr(gemm_umma2_kernel<EPI>, GEMM2_SMEM_BYTES, attr_done)) return 1;
  int sms = device_sm_count();
  if (sms <= 0) return fail("no CUDA device");
  int clusters = sms / 2;
  if (prm.total_tiles < clusters) clusters = prm.total_tiles;
  gemm_umma2_kernel<EPI><<<2 * clusters, GEMM_THREADS, GEMM2_SMEM_BYTES, st>>>(*tA[0], *tB[0], *tA[1], *tB[1], prm);
  QIMG_LAUNCH_CHECK("gemm_umma2_kernel");
  return 0;
}

// 0 = one CTA per tile (128x256, cta_group::1); 1 = CTA pair per tile (256x256, cta_group::2)
static int g_gemm_auto_tile = 1;  // qimg_set_gemm_mode(2): pair tiles unless 128-row tiles save whole rounds (default)
static int g_gemm_mode = -1;
static int gemm_mode() {
  if (g_gemm_mode < 0) {
    const char* e = getenv("QIMG_GEMM_MODE");
    const int m = e ? atoi(e) : 2;  // CTA pair: +6 % over cta_group::1 on every block GEMM (profiles/r01_kernel_bench.md)
    g_gemm_mode = m == 0 ? 0 : 1;
    g_gemm_auto_tile = (m == 2) ? 1 : 0;
  }
  return g_gemm_mode;
}

// raster band height of the CTA-pair kernel in 256-row tiles (the 1-CTA kernel uses twice as many 128-row tiles): the
// tiles of a band sweep every weight column while the band's activations stay L2-resident; each band re-streams the weight
// matrix from HBM once, so taller bands cut weight re-reads (8 bands -> 4 at M = 16384 with 16)
static int g_gemm_group_m = -1;
static int gemm_group_m() {
  if (g_gemm_group_m < 0) {
    const char* e = getenv("QIMG_GEMM_GROUP_M");
    g_gemm_group_m = e ? atoi(e) : GEMM2_GROUP_M;
    if (g_gemm_group_m < 1) g_gemm_group_m = GEMM2_GROUP_M;
  }
  return g_gemm_group_m;
}

static int launch_gemm(const qimg_gemm_problem* pr, int nprob, int epi, cudaStream_t st) {
  if (nprob < 1 || nprob > 2) return fail("qimg_gemm: nprob must be 1 or 2");
  tmap_cache_trim();
  int maxN = 0;
  for (int i = 0; i < nprob; ++i) maxN = pr[i].N > maxN ? pr[i].N : maxN;
  const int BN = (maxN <= 64 && epi == QIMG_EPI_BIAS) ? 64 : 256;
  // tile shape: the CTA pair (256 x 256, cta_group::2) is 6 % faster per FLOP, but small M (one image under sequence /
  // tensor parallelism: M = 1088 is 4.25 pair tiles) loses whole rounds to tile quantisation; 128-row tiles on single CTAs
  // then need fewer rounds.  Estimated cost = rounds x relative tile time; both kernels run the same K loop per output
  // element, so the choice does not change a single bit of the result (test_gemm_tile_modes_bit_identical).
  bool pair = (BN == 256) && gemm_mode() == 1;
  if (pair && g_gemm_auto_tile) {
    const int sms = device_sm_count();
    long long t2 = 0, t1 = 0;
    for (int i = 0; i < nprob; ++i) {
      const long long nt = (pr[i].N + 255) / 256;
      t2 += (long long)((pr[i].M + 255) / 256) * nt;
      t1 += (long long)((pr[i].M + 127) / 128) * nt;
    }
    if (sms >= 2) {
      const long long r2 = (t2 + sms / 2 - 1) / (sms / 2), r1 = (t1 + sms - 1) / sms;
      if ((double)r1 * 1.06 < (double)r2) pair = false;
    }
  }
  const int tile_m = pair ? 256 : GEMM_BM;
  GemmParams prm;
  memset(&prm, 0, sizeof prm);
  prm.nprob = nprob;
  const CUtensorMap* tA[2] = {nullptr, nullptr};
  const CUtensorMap* tB[2] = {nullptr, nullptr};
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    const qimg_gemm_problem& s = pr[i];
    if (s.M <= 0 || s.N <= 0 || s.K <= 0) return fail("qimg_gemm: empty problem");
    if (s.K % 8 || s.N % 8) return fail("qimg_gemm: K and N must be multiples of 8");
    if (s.rows_per_batch <= 0) return fail("qimg_gemm: rows_per_batch must be > 0");
    GemmProblem& d = prm.p[i];
    d.M = s.M; d.N = s.N; d.K = s.K;
    d.rows_per_batch = s.rows_per_batch;
    d.bias = (const bf16*)s.bias;
    d.out = (bf16*)s.out;
    d.ldo = s.ldo;
    d.gate = (const bf16*)s.gate;
    d.gate_stride = s.gate_stride;
    d.q = (bf16*)s.q; d.k = (bf16*)s.k; d.v = (bf16*)s.v;
    d.nq_w = (const bf16*)s.norm_q_w; d.nk_w = (const bf16*)s.norm_k_w;
    d.cos = (const bf16*)s.rope_cos; d.sin = (const bf16*)s.rope_sin;
    d.S_joint = s.S_joint; d.pos_off = s.pos_off; d.H = s.H; d.eps = s.eps;
    if (s.row_base < 0) return fail("qimg_gemm: negative row_base");
    d.row_base = s.row_base;
    d.sp_size = s.sp_size > 1 ? s.sp_size : 1;
    if (d.sp_size > 1) {
      if (epi != QIMG_EPI_QKV) return fail("qimg_gemm: sp_size > 1 is only meaningful with the QKV epilogue");
      if (d.sp_size > 8 || s.H % d.sp_size) return fail("qimg_gemm: sp_size must divide H and be <= 8");
      for (int r = 0; r < d.sp_size; ++r) {
        if (!s.sp_q[r] || !s.sp_k[r] || !s.sp_v[r]) return fail("qimg_gemm: null sp_q / sp_k / sp_v pointer");
        d.sp_q[r] = (bf16*)s.sp_q[r]; d.sp_k[r] = (bf16*)s.sp_k[r]; d.sp_v[r] = (bf16*)s.sp_v[r];
      }
    }
    if (epi == QIMG_EPI_PARTIAL_F32) {
      if (s.tp_size < 1 || s.tp_size > 8 || s.tp_rank < 0 || s.tp_rank >= s.tp_size)
        return fail("qimg_gemm: partial-sum epilogue needs 1 <= tp_size <= 8 and 0 <= tp_rank < tp_size");
      if (s.tp_recv_rows <= 0 || s.tp_recv_row_off < 0) return fail("qimg_gemm: bad tp_recv_rows / tp_recv_row_off");
      for (int r = 0; r < s.tp_size; ++r) {
        if (!s.tp_recv[r]) return fail("qimg_gemm: null tp_recv pointer");
        d.tp_recv[r] = (float*)s.tp_recv[r];
      }
      d.tp_size = s.tp_size; d.tp_rank = s.tp_rank; d.tp_recv_rows = s.tp_recv_rows; d.tp_recv_row_off = s.tp_recv_row_off;
    } else if (!s.bias) {
      return fail("qimg_gemm: bias is required");
    }
    if (epi == QIMG_EPI_QKV) {
      if (s.N != 3 * s.H * 128) return fail("qimg_gemm: QKV epilogue needs N == 3*H*128");
      if ((d.sp_size == 1 && (!s.q || !s.k || !s.v)) || !s.norm_q_w || !s.norm_k_w || !s.rope_cos || !s.rope_sin)
        return fail("qimg_gemm: QKV epilogue pointers missing");
    } else if (epi != QIMG_EPI_PARTIAL_F32 && (!s.out || s.ldo % 8)) {
      return fail("qimg_gemm: out missing or ldo not a multiple of 8");
    }
    if (epi == QIMG_EPI_BIAS_GATE_RES && !s.gate) return fail("qimg_gemm: gate missing");
    d.m_tiles = (s.M + tile_m - 1) / tile_m;
    d.n_tiles = (s.N + BN - 1) / BN;
    d.tile_begin = tiles;
    tiles += d.m_tiles * d.n_tiles;
    tA[i] = get_tmap_2d(s.A, (uint64_t)s.K, (uint64_t)s.M, GEMM_BM);
    tB[i] = get_tmap_2d(s.W, (uint64_t)s.K, (uint64_t)s.N, pair ? 128u : (uint32_t)BN);
    if (!tA[i] || !tB[i]) return 1;
  }
  if (nprob == 1) {
    tA[1] = tA[0];
    tB[1] = tB[0];
  }
  prm.total_tiles = tiles;
  prm.skip = launch_predicate();
  prm.group_m = pair ? gemm_group_m() : 2 * gemm_group_m();
  double flops = 0;
  for (int i = 0; i < nprob; ++i) flops += 2.0 * pr[i].M * (double)pr[i].N * pr[i].K;
  ProfScope prof(0, flops, st);
  if (BN == 64) return launch_gemm_inst<64, EPI_BIAS>(tA, tB, prm, st);
  if (pair) {
    switch (epi) {
      case QIMG_EPI_BIAS: return launch_gemm2_inst<EPI_BIAS>(tA, tB, prm, st);
      case QIMG_EPI_BIAS_GELU: return launch_gemm2_inst<EPI_BIAS_GELU>(tA, tB, prm, st);
      case QIMG_EPI_BIAS_GATE_RES: return launch_gemm2_inst<EPI_BIAS_GATE_RES>(tA, tB, prm, st);
      case QIMG_EPI_QKV: return launch_gemm2_inst<EPI_QKV>(tA, tB, prm, st);
      case QIMG_EPI_PARTIAL_F32: return launch_gemm2_inst<EPI_PARTIAL_F32>(tA, tB, prm, st);
    }
    return fail("qimg_gemm: unknown epilogue");
  }
  switch (epi) {
    case QIMG_EPI_BIAS: return launch_gemm_inst<256, EPI_BIAS>(tA, tB, prm, st);
    case QIMG_EPI_BIAS_GELU: return launch_gemm_inst<256, EPI_BIAS_GELU>(tA, tB, prm, st);
    case QIMG_EPI_BIAS_GATE_RES: return launch_gemm_inst<256, EPI_BIAS_GATE_RES>(tA, tB, prm, st);
    case QIMG_EPI_QKV: return launch_gemm_inst<256, EPI_QKV>(tA, tB, prm, st);
    case QIMG_EPI_PARTIAL_F32: return launch_gemm_inst<256, EPI_PARTIAL_F32>(tA, tB, prm, st);
  }
  return fail("qimg_gemm: unknown epilogue");
}

// ------------------------------------------------------------------------------------------
// raw tcgen05 probe (test tool): one CTA, D[128,128] = A[128,128] * B^T, three operand paths
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(192, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const bf16* __restrict__ Ag, float* __restrict__ Dg, int mode) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* ld_full = bars;
  uint64_t* mma_done = bars + 1;
  uint64_t* p_ready = bars + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    mbar_init(ld_full, 1);
    mbar_init(mma_done, 1);
    mbar_init(p_ready, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t tD = tmem_base, tP = tmem_base + 128;
  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(ld_full, 65536);
      for (int s = 0; s < 2; ++s) {
        tma_load_2d(sA + s * 16384, &tmA, ld_full, s * 64, 0);
        tma_load_2d(sB + s * 16384, &tmB, ld_full, s * 64, 0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(ld_full, 0);
      if (mode == 2) mbar_wait(p_ready, 0);
      tc_fence_after();
      const uint32_t a = smem_u32(sA), b = smem_u32(sB);
      for (int k = 0; k < 8; ++k) {
        const uint32_t koff = (k >> 2) * 16384 + (k & 3) * 32;
        if (mode == 0)
          umma_ss(tD, make_kmajor_sw128_desc(a + koff), make_kmajor_sw128_desc(b + koff), make_idesc_bf16(128, 128, 0, 0), k != 0);
        else if (mode == 1)
          umma_ss(tD, make_kmajor_sw128_desc(a + koff), make_mnmajor_sw128_desc(b + k * 2048, 16384),
                  make_idesc_bf16(128, 128, 0, 1), k != 0);
        else
          umma_ts(tD, tP + k * 8, make_mnmajor_sw128_desc(b + k * 2048, 16384), make_idesc_bf16(128, 128, 0, 1), k != 0);
      }
      umma_commit(mma_done);
    }
  } else {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    if (mode == 2) {
      // stage A row-per-lane into TMEM as packed bf16 pairs (the FMHA "P" operand path)
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t pk[16];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(Ag + (size_t)row * 128 + cc * 32);
        for (int i = 0; i < 16; ++i) pk[i] = src[i];
        tmem_st_32x32b_x16(tP + lane_off + cc * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    mbar_wait(mma_done, 0);
    tc_fence_after();
    for (int cc = 0; cc < 4; ++cc) {
      uint32_t r[32];
      tmem_ld_32x32b_x32(tD + lane_off + cc * 32, r);
      tmem_ld_wait();
      for (int i = 0; i < 32; ++i) Dg[(size_t)row * 128 + cc * 32 + i] = __uint_as_float(r[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace qimg

static long long* g_fmha_trace = nullptr;  // qimg_set_fmha_trace (diagnostics)
using namespace qimg;

// Overflow flag of the fast attention pipeline: one int per device, owned by the library (4 bytes, allocated on first use)
static int* g_fmha_ovf[qimg::kMaxDevices] = {};
static int* fmha_overflow_flag() {
  const int dev = current_device();
  if (dev < 0) return nullptr;
  if (!g_fmha_ovf[dev]) {
    if (cudaMalloc(&g_fmha_ovf[dev], sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(g_fmha_ovf[dev], 0, sizeof(int));
  }
  return g_fmha_ovf[dev];
}

template <uint32_t MASK>
static int launch_fmha_inst(int pipeline, const CUtensorMap* tq, const CUtensorMap* tk, const CUtensorMap* tv,
                            const FmhaParams& prm, cudaStream_t st) {
  static bool done7[kMaxDevices] = {}, done9[kMaxDevices] = {};

  const int pairs = (prm.S + 255) / 256;
  const int grid = prm.single_tile ? ((prm.S + 127) / 128) * prm.B * prm.H : pairs * prm.B * prm.H;
  if (pipeline == 6) {
    if (ensure_smem_attr(fmha_joint_kernel_v9<MASK>, FMHA4_SMEM_BYTES, done9)) return 1;
    fmha_joint_kernel_v9<MASK><<<grid, FMHA4_THREADS, FMHA4_SMEM_BYTES, st>>>(*tq, *tk, *tv, prm);
  } else {
    if (ensure_smem_attr(fmha_joint_kernel_v7<MASK>, FMHA4_SMEM_BYTES, done7)) return 1;
    fmha_joint_kernel_v7<MASK><<<grid, FMHA4_THREADS, FMHA4_SMEM_BYTES, st>>>(*tq, *tk, *tv, prm);
  }
  QIMG_LAUNCH_CHECK("fmha_joint_kernel");
  return 0;
}

extern "C" {

int qimg_abi_version(void) { return 1; }
const char* qimg_last_error(void) { return g_last_error.c_str(); }
long long qimg_launch_count(void) { return g_launch_count.load(); }
void qimg_reset_launch_count(void) { g_launch_count.store(0); }

void qimg_prof_enable(int on) { g_prof_on.store(on != 0); }

int qimg_set_gemm_mode(int mode) {
  if (mode < 0 || mode > 2) return fail("qimg_set_gemm_mode: mode must be 0 (cta_group::1), 1 (cta_group::2 pair) or 2 (pair unless single saves rounds)");
  g_gemm_mode = mode == 0 ? 0 : 1;
  g_gemm_auto_tile = mode == 2;
  return 0;
}
int qimg_get_gemm_mode(void) { return gemm_mode() == 0 ? 0 : (g_gemm_auto_tile ? 2 : 1); }
int qimg_set_gemm_group_m(int tiles) {
  if (tiles < 1 || tiles > 1024) return fail("qimg_set_gemm_group_m: band height must be in [1, 1024] tiles");
  g_gemm_group_m = tiles;
  return 0;
}

int qimg_prof_collect(int kind, double* ms_total, long long* launches, double* flops_total) {
  if (kind < 0 || kind > 1) return fail("qimg_prof_collect: kind");
  double ms = 0, fl = 0;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (ProfRec& r : g_prof[kind]) {
    QIMG_CUDA_CHECK(cudaEventSynchronize(r.b));
    float t = 0;
    QIMG_CUDA_CHECK(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    fl += r.flops;
    g_event_pool.push_back(r.a);
    g_event_pool.push_back(r.b);
  }
  if (ms_total) *ms_total = ms;
  if (launches) *launches = (long long)g_prof[kind].size();
  if (flops_total) *flops_total = fl;
  g_prof[kind].clear();
  return 0;
}

int qimg_device_check(int* sm_count) {
  int dev = 0, major = 0, minor = 0, n = 0;
  QIMG_CUDA_CHECK(cudaGetDevice(&dev));
  QIMG_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  QIMG_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  QIMG_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (sm_count) *sm_count = n;
  if (major != 10) {
    char buf[64];
    snprintf(buf, sizeof buf, "device is sm_%d%d", major, minor);
    return fail("qimg_b200 requires an sm_100 (B200) device", buf);
  }
  return 0;
}

static int ln_modulate_launch(const void* x, const void* shift, const void* scale, void* y, int rows, int row_base, int D,
                              int rows_per_batch, long long mod_stride, float eps, const int* index, int index_batch,
                              qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8 || D > EW_MAX_CHUNKS * 256) return fail("qimg_ln_modulate: D must be a multiple of 8 and <= 4096");
  if (rows_per_batch <= 0 || row_base < 0) return fail("qimg_ln_modulate: rows_per_batch / row_base");
  const dim3 grid((rows + 3) / 4);
  cudaStream_t st = (cudaStream_t)stream;
  const bf16 *xp = (const bf16*)x, *shp = (const bf16*)shift, *scp = (const bf16*)scale;
  const int* skip = launch_predicate();
  if (D == 3072) {  // Qwen-Image width: compile-time row length
    ln_modulate_fast_kernel<12><<<grid, 128, 0, st>>>(xp, shp, scp, (bf16*)y, rows, row_base, rows_per_batch, mod_stride, eps, skip,
                                                      index, index_batch);
  } else if (D == 1024) {
    ln_modulate_fast_kernel<4><<<grid, 128, 0, st>>>(xp, shp, scp, (bf16*)y, rows, row_base, rows_per_batch, mod_stride, eps, skip,
                                                     index, index_batch);
  } else if (D == 256) {
    ln_modulate_fast_kernel<1><<<grid, 128, 0, st>>>(xp, shp, scp, (bf16*)y, rows, row_base, rows_per_batch, mod_stride, eps, skip,
                                                     index, index_batch);
  } else {
    ln_modulate_kernel<<<grid, 128, 0, st>>>(xp, shp, scp, (bf16*)y, rows, row_base, D, rows_per_batch, mod_stride, eps, skip, index,
                                             index_batch);
  }
  QIMG_LAUNCH_CHECK("ln_modulate_kernel");
  return 0;
}

int qimg_ln_modulate_rows(const void* x, const void* shift, const void* scale, void* y, int rows, int row_base, int D,
                          int rows_per_batch, long long mod_stride, float eps, qimg_stream_t stream) {
  return ln_modulate_launch(x, shift, scale, y, rows, row_base, D, rows_per_batch, mod_stride, eps, nullptr, 0, stream);
}

int qimg_ln_modulate(const void* x, const void* shift, const void* scale, void* y, int rows, int D, int rows_per_batch,
                     long long mod_stride, float eps, qimg_stream_t stream) {
  return ln_modulate_launch(x, shift, scale, y, rows, 0, D, rows_per_batch, mod_stride, eps, nullptr, 0, stream);
}

int qimg_ln_modulate_indexed(const void* x, const void* shift, const void* scale, void* y, int rows, int D, int rows_per_batch,
                             long long mod_stride, float eps, const int* index, int index_batch, qimg_stream_t stream) {
  if (!index || index_batch < 1) return fail("qimg_ln_modulate_indexed: index [rows] (int32, device) and the batch size are required");
  return ln_modulate_launch(x, shift, scale, y, rows, 0, D, rows_per_batch, mod_stride, eps, index, index_batch, stream);
}

int qimg_select_rows(const void* src, long long src_stride, const int* index, void* out, int rows, int D, int rows_per_batch,
                     int index_batch, qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (!src || !index || !out || D % 8 || rows_per_batch <= 0 || index_batch < 1) return fail("qimg_select_rows: bad arguments (D % 8 == 0)");
  const long long n_vec = (long long)rows * (D / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  select_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)src, src_stride, index, (bf16*)out, n_vec, D,
                                                                    rows_per_batch, index_batch);
  QIMG_LAUNCH_CHECK("select_rows_kernel");
  return 0;
}

int qimg_rms_norm(const void* x, const void* w, void* y, int rows, int D, float eps, qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8 || D > EW_MAX_CHUNKS * 256) return fail("qimg_rms_norm: D must be a multiple of 8 and <= 4096");
  rms_norm_kernel<<<(rows + 3) / 4, 128, 0, (cudaStream_t)stream>>>((const bf16*)x, (const bf16*)w, (bf16*)y, rows, D, eps);
  QIMG_LAUNCH_CHECK("rms_norm_kernel");
  return 0;
}

int qimg_gate_residual(void* x, const void* y, const void* gate, int rows, int D, int rows_per_batch,
                       long long gate_stride, qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8) return fail("qimg_gate_residual: D must be a multiple of 8");
  const long long n_vec = (long long)rows * (D / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gate_residual_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)x, (const bf16*)y, (const bf16*)gate, n_vec, D,
                                                                      rows_per_batch, gate_stride);
  QIMG_LAUNCH_CHECK("gate_residual_kernel");
  return 0;
}

int qimg_gate_residual_bias(void* x, const void* y, const void* bias, const void* gate, int rows, int D, int rows_per_batch,
                            long long gate_stride, qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (D % 8) return fail("qimg_gate_residual_bias: D must be a multiple of 8");
  const long long n_vec = (long long)rows * (D / 8);
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  gate_residual_bias_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((bf16*)x, (const bf16*)y, (const bf16*)bias,
                                                                           (const bf16*)gate, n_vec, D, rows_per_batch, gate_stride);
  QIMG_LAUNCH_CHECK("gate_residual_bias_kernel");
  return 0;
}

static int ew_grid(long long n_vec) {
  long long blocks = (n_vec + 255) / 256;
  const long long cap = (long long)device_sm_count() * 16;
  return (int)(blocks > cap ? cap : blocks);
}

int qimg_rel_l1_sums(const void* a, const void* b, long long n, float* sums2, qimg_stream_t stream) {
  if (n <= 0) return fail("qimg_rel_l1_sums: empty input");
  if (n % 8) return fail("qimg_rel_l1_sums: n must be a multiple of 8");
  QIMG_CUDA_CHECK(cudaMemsetAsync(sums2, 0, 2 * sizeof(float), (cudaStream_t)stream));
  rel_l1_sums_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((const bf16*)a, (const bf16*)b, n / 8, sums2);
  QIMG_LAUNCH_CHECK("rel_l1_sums_kernel");
  return 0;
}

int qimg_bf16_sub(void* out, const void* a, const void* b, long long n, qimg_stream_t stream) {
  if (n <= 0) return 0;
  if (n % 8) return fail("qimg_bf16_sub: n must be a multiple of 8");
  bf16_sub_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((bf16*)out, (const bf16*)a, (const bf16*)b, n / 8);
  QIMG_LAUNCH_CHECK("bf16_sub_kernel");
  return 0;
}

int qimg_tea_decide(const float* sums2, long long n, const double* coef5, double thresh, double* accum, int* flag, float* hist,
                    int hist_idx, int force, qimg_stream_t stream) {
  if (!accum || !flag || !coef5) return fail("qimg_tea_decide: null argument");
  if (force == 0 && (!sums2 || n <= 0)) return fail("qimg_tea_decide: sums / n missing");
  TeaCoef c;
  for (int i = 0; i < 5; ++i) c.c[i] = coef5[i];
  tea_decide_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(sums2, (double)n, c, thresh, accum, flag, hist, hist_idx, force);
  QIMG_LAUNCH_CHECK("tea_decide_kernel");
  return 0;
}

int qimg_tea_residual(void* x, const void* ori, void* resid, long long n, const int* flag, qimg_stream_t stream) {
  if (n <= 0) return 0;
  if (n % 8) return fail("qimg_tea_residual: n must be a multiple of 8");
  if (!flag) return fail("qimg_tea_residual: null flag");
  tea_residual_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((bf16*)x, (const bf16*)ori, (bf16*)resid, n / 8, flag);
  QIMG_LAUNCH_CHECK("tea_residual_kernel");
  return 0;
}

int qimg_bf16_add_inplace(void* x, const void* r, long long n, qimg_stream_t stream) {
  if (n <= 0) return 0;
  if (n % 8) return fail("qimg_bf16_add_inplace: n must be a multiple of 8");
  bf16_add_inplace_kernel<<<ew_grid(n / 8), 256, 0, (cudaStream_t)stream>>>((bf16*)x, (const bf16*)r, n / 8);
  QIMG_LAUNCH_CHECK("bf16_add_inplace_kernel");
  return 0;
}

int qimg_linear_small_m(const void* x, const void* W, const void* bias, void* y, int M, long long N, int K,
                        long long ldy, int act_silu, qimg_stream_t stream) {
  if (M <= 0 || N <= 0) return 0;
  if (K % 8) return fail("qimg_linear_small_m: K must be a multiple of 8");
  if (M > 64) return fail("qimg_linear_small_m: M > 64 (use qimg_gemm)");
  static bool attr_done[kMaxDevices] = {};
  if (ensure_smem_attr(linear_small_m_kernel<8>, 8 * 4096 * 2, attr_done)) return 1;
  if (K > 4096) return fail("qimg_linear_small_m: K > 4096");
  const int sms = device_sm_count();
  for (int m0 = 0; m0 < M; m0 += 8) {
    const int mm = (M - m0) < 8 ? (M - m0) : 8;
    long long blocks = (N + 7) / 8;
    const long long cap = (long long)sms * 8;
    if (blocks > cap) blocks = cap;
    linear_small_m_kernel<8><<<(int)blocks, 256, (size_t)mm * K * 2, (cudaStream_t)stream>>>(
        (const bf16*)x + (size_t)m0 * K, (const bf16*)W, (const bf16*)bias, (bf16*)y + (size_t)m0 * ldy, mm, N, K, ldy, act_silu);
    QIMG_LAUNCH_CHECK("linear_small_m_kernel");
  }
  return 0;
}

int qimg_timestep_sinusoid(const void* t, void* out, int B, qimg_stream_t stream) {
  if (B <= 0) return 0;
  timestep_sinusoid_kernel<<<(B * 256 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const bf16*)t, (bf16*)out, B);
  QIMG_LAUNCH_CHECK("timestep_sinusoid_kernel");
  return 0;
}

static int g_euler_dt_fp32 = 0;
int qimg_set_euler_dt_fp32(int on) {
  g_euler_dt_fp32 = on != 0;
  return 0;
}

static int launch_cfg_euler(const void* pos, const void* neg, void* latents, long long rows, int C, float cfg_scale, float dt,
                            const float* sigma_pair, qimg_stream_t stream) {
  if (rows <= 0) return 0;
  if (C != 64) return fail("qimg_cfg_euler_step: C must be 64 (packed latent channels)");
  const long long vecs = rows * 8;
  cfg_euler_step_kernel<<<(int)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)pos, (const bf16*)neg,
                                                                                     (bf16*)latents, rows, cfg_scale, dt,
                                                                                     sigma_pair, g_euler_dt_fp32);
  QIMG_LAUNCH_CHECK("cfg_euler_step_kernel");
  return 0;
}

int qimg_cfg_euler_step(const void* pos, const void* neg, void* latents, long long rows, int C, float cfg_scale,
                        float sigma, float sigma_next, qimg_stream_t stream) {
  // fp32 subtraction, as the scheduler's 0-dim fp32 tensors
  return launch_cfg_euler(pos, neg, latents, rows, C, cfg_scale, sigma_next - sigma, nullptr, stream);
}

int qimg_cfg_euler_step_dev(const void* pos, const void* neg, void* latents, long long rows, int C, float cfg_scale,
                            const float* sigma_pair, qimg_stream_t stream) {
  if (!sigma_pair) return fail("qimg_cfg_euler_step_dev: null sigma pair");
  return launch_cfg_euler(pos, neg, latents, rows, C, cfg_scale, 0.f, sigma_pair, stream);
}

int qimg_gemm(const qimg_gemm_problem* problems, int nprob, int epilogue, qimg_stream_t stream) {
  if (!problems) return fail("qimg_gemm: null problems");
  return launch_gemm(problems, nprob, epilogue, (cudaStream_t)stream);
}

// mode = pipeline | (poly << 3): pipeline 6 = fast (delayed reference maximum, guarded by the overflow flag; DEFAULT),
// pipeline 4 = exact (every tile's maximum reduced first); poly 1 = 25 % of the exponentials on the FMA-pipe polynomial
static int g_fmha_single_tile = -1;  // -1: auto (one query tile per CTA while that fits one wave)
static int g_fmha_mode = -1;
static int fmha_mode() {
  if (g_fmha_mode < 0) {
    const char* e = getenv("QIMG_FMHA_MODE");
    g_fmha_mode = e ? atoi(e) : 6;
    if ((g_fmha_mode & 7) != 4 && (g_fmha_mode & 7) != 6) g_fmha_mode = 6;
  }
  return g_fmha_mode;
}
int qimg_set_fmha_mode(int mode) {
  if (mode < 0 || mode > 15 || ((mode & 7) != 4 && (mode & 7) != 6))
    return fail("qimg_set_fmha_mode: mode must be 4 (exact) or 6 (fast), optionally | 8 (25 % polynomial exponentials)");
  g_fmha_mode = mode;
  return 0;
}
int qimg_get_fmha_mode(void) { return fmha_mode(); }
int qimg_set_fmha_single_tile(int mode) {
  if (mode < -1 || mode > 1) return fail("qimg_set_fmha_single_tile: -1 (auto), 0 (query-tile pairs) or 1 (one query tile per CTA)");
  g_fmha_single_tile = mode;
  return 0;
}

int qimg_set_fmha_trace(void* dev_buf_32_i64) {
  if (dev_buf_32_i64 && !kFmhaTrace) return fail("qimg_set_fmha_trace: library built without -DQIMG_FMHA_TRACE");
  g_fmha_trace = (long long*)dev_buf_32_i64;
  return 0;
}

int qimg_fmha_overflow(int* out, int reset) {
  int* flag = fmha_overflow_flag();
  if (!flag) return fail("qimg_fmha_overflow: no CUDA device / allocation failed");
  int v = 0;
  QIMG_CUDA_CHECK(cudaMemcpy(&v, flag, sizeof(int), cudaMemcpyDeviceToHost));  // synchronises with prior launches
  if (reset && v) QIMG_CUDA_CHECK(cudaMemset(flag, 0, sizeof(int)));
  if (out) *out = v;
  return 0;
}

static int fmha_launch(const void* q, const void* k, const void* v, void* out_txt, void* out_img, int B, int H, int S, int T,
                       float softmax_scale, int mode, const qimg_fmha_sp* sp, qimg_stream_t stream) {
  if (B <= 0 || H <= 0 || S <= 0 || T < 0 || T > S) return fail("qimg_fmha_joint: bad shape");
  if (mode < 0) mode = fmha_mode();
  const int pipeline = mode & 7;
  if (mode > 15 || (pipeline != 4 && pipeline != 6)) return fail("qimg_fmha_joint: mode must be 4 (exact) or 6 (fast) [| 8]");
  tmap_cache_trim();
  const CUtensorMap* tq = get_tmap_3d(q, 128, (uint64_t)S, (uint64_t)B * H, 128);
  const CUtensorMap* tk = get_tmap_3d(k, 128, (uint64_t)S, (uint64_t)B * H, 128);
  const CUtensorMap* tv = get_tmap_3d(v, 128, (uint64_t)S, (uint64_t)B * H, 128);
  if (!tq || !tk || !tv) return 1;
  FmhaParams prm;
  memset(&prm, 0, sizeof prm);
  prm.out_txt = (bf16*)out_txt;
  prm.out_img = (bf16*)out_img;
  prm.B = B; prm.H = H; prm.S = S; prm.T = T;
  prm.scale_log2 = softmax_scale * 1.4426950408889634f;
  prm.trace = g_fmha_trace;
  prm.sp_size = 1;
  if (sp && sp->sp_size > 1) {
    if (sp->sp_size > 8 || sp->sp_rank < 0 || sp->sp_rank >= sp->sp_size) return fail("qimg_fmha_joint_sp: bad sp_size / sp_rank");
    prm.sp_size = sp->sp_size; prm.sp_rank = sp->sp_rank;
    prm.sp_rows_img = B * (S - T); prm.sp_rows_txt = B * T;
    for (int r = 0; r < sp->sp_size; ++r) {
      if (!sp->out_img[r] || (T > 0 && !sp->out_txt[r])) return fail("qimg_fmha_joint_sp: null owner buffer");
      prm.sp_out_img[r] = (bf16*)sp->out_img[r];
      prm.sp_out_txt[r] = (bf16*)sp->out_txt[r];
    }
  } else if (!out_img || (T > 0 && !out_txt)) {
    return fail("qimg_fmha_joint: null output");
  }
  prm.skip = launch_predicate();
  {
    const int sms = device_sm_count();
    // one query tile per CTA only while that grid still fits ONE wave (e.g. the 3 local heads of an 8-way split at B = 1:
    // 99 CTAs instead of 51): a one-tile CTA has no second tile whose MMAs could run under its softmax, so it takes ~0.8 of a
    // pair CTA's time, not 0.5 — with more than one wave the pair grid wins (a "rounds x 0.55" rule chose one-tile CTAs at
    // S = 1152, B = 4 and ran at 523 instead of 703 TFLOP/s, profiles/r02_fmha_sweep_4.log).  Results do not depend on it.
    const int single = (sms > 0 && (long long)((S + 127) / 128) * B * H <= sms) ? 1 : 0;
    prm.single_tile = g_fmha_single_tile >= 0 ? g_fmha_single_tile : single;
  }
  prm.overflow = fmha_overflow_flag();
  if (!prm.overflow) return fail("qimg_fmha_joint: could not allocate the overflow flag");
  ProfScope prof(1, 4.0 * B * H * (double)S * S * 128, (cudaStream_t)stream);
  if (mode & 8) return launch_fmha_inst<0x11u>(pipeline, tq, tk, tv, prm, (cudaStream_t)stream);
  return launch_fmha_inst<0x00u>(pipeline, tq, tk, tv, prm, (cudaStream_t)stream);
}

int qimg_fmha_joint_mode(const void* q, const void* k, const void* v, void* out_txt, void* out_img, int B, int H, int S,
                         int T, float softmax_scale, int mode, qimg_stream_t stream) {
  return fmha_launch(q, k, v, out_txt, out_img, B, H, S, T, softmax_scale, mode, nullptr, stream);
}

int qimg_fmha_joint_sp(const void* q, const void* k, const void* v, int B, int H_local, int S, int T, float softmax_scale,
                       const qimg_fmha_sp* sp, qimg_stream_t stream) {
  if (!sp) return fail("qimg_fmha_joint_sp: null descriptor");
  return fmha_launch(q, k, v, nullptr, nullptr, B, H_local, S, T, softmax_scale, -1, sp, stream);
}

int qimg_fmha_joint(const void* q, const void* k, const void* v, void* out_txt, void* out_img, int B, int H, int S,
                    int T, float softmax_scale, qimg_stream_t stream) {
  return qimg_fmha_joint_mode(q, k, v, out_txt, out_img, B, H, S, T, softmax_scale, -1, stream);
}

int qimg_umma_probe(const void* A, const void* B, float* D, int N, int K, int mode, qimg_stream_t stream) {
  if (N != 128 || K != 128) return fail("qimg_umma_probe: N and K must be 128");
  if (mode < 0 || mode > 2) return fail("qimg_umma_probe: mode");
  // A [128 rows, K] K-major.  B: mode 0 -> [N, K] (K contiguous); modes 1,2 -> [K, N] (N contiguous).
  const CUtensorMap* ta = get_tmap_2d(A, 128, 128, 128);
  const CUtensorMap* tb = get_tmap_2d(B, 128, 128, 128);
  if (!ta || !tb) return 1;
  static bool attr_done[kMaxDevices] = {};
  if (ensure_smem_attr(umma_probe_kernel, 65536 + 1024 + 256, attr_done)) return 1;
  umma_probe_kernel<<<1, 192, 65536 + 1024 + 256, (cudaStream_t)stream>>>(*ta, *tb, (const bf16*)A, D, mode);
  QIMG_LAUNCH_CHECK("umma_probe_kernel");
  return 0;
}

}  // extern "C"
