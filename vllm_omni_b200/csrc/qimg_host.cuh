// Host-side plumbing shared by the C-ABI translation units: error reporting, launch counting,
// driver entry point for cuTensorMapEncodeTiled (no link-time libcuda dependency) and a
// process-wide TMA descriptor cache keyed by (pointer, shape, box).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>

namespace qimg {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launch_count;

inline int fail(const char* what, const char* detail = nullptr) {
  g_last_error = what;
  if (detail) {
    g_last_error += ": ";
    g_last_error += detail;
  }
  return 1;
}

#define QIMG_CUDA_CHECK(expr)                                     \
  do {                                                            \
    cudaError_t _e = (expr);                                      \
    if (_e != cudaSuccess) return fail(#expr, cudaGetErrorString(_e)); \
  } while (0)

#define QIMG_LAUNCH_CHECK(name)                                    \
  do {                                                             \
    cudaError_t _e = cudaGetLastError();                           \
    if (_e != cudaSuccess) return fail(name, cudaGetErrorString(_e)); \
    g_launch_count.fetch_add(1, std::memory_order_relaxed);        \
  } while (0)

int device_sm_count();  // cached; <= 0 on error

// Device predicate applied to the launches of the CURRENT host thread (LN-modulate, tcgen05 GEMM, attention): while set,
// every such kernel starts with `if (*flag) return;`.  The engine brackets its BLOCKS stage with it when a step cache
// decided on the device (qimg_engine_set_blocks_predicate); nullptr = unconditional launches.
void set_launch_predicate(const int* flag);
const int* launch_predicate();

// bf16 row-major matrix view [dim1 = rows][dim0 = cols] (optionally x dim2 batches), SWIZZLE_128B,
// inner box = 64 elements (128 B).  Returns nullptr on failure (error recorded).
const CUtensorMap* get_tmap_2d(const void* ptr, uint64_t cols, uint64_t rows, uint32_t box_rows);
void tmap_cache_trim();  // drop the descriptor cache when it has grown large; call only at the start of a launcher
const CUtensorMap* get_tmap_3d(const void* ptr, uint64_t cols, uint64_t rows, uint64_t batches, uint32_t box_rows);

}  // namespace qimg
