// Host-side interface between the engine (qimg_engine.cu) and the peer-memory tensor-parallel kernels (qimg_tp_p2p.cu).
#pragma once

#include <cuda_runtime.h>

namespace qimg {

// One reduction of the sequence-parallel TP epilogue over both streams (index 0 = image rows, 1 = text rows).
struct TpReduceArgs {
  int P, rank, D;
  const void* recv_local;  // this rank's receive buffer [P][recv_rows][D] fp32
  int recv_rows;           // rows per source rank (image slice + text slice, each rounded up to the largest owner slice)
  int row_off[2];          // first receive row of each stream within a source's block
  void* x[2];              // local residual streams [rows, D] bf16
  int rows[2], rows_per_batch[2];
  const void* bias[2];
  const void* gate[2];
  const void* shift[2];    // modulation of the LayerNorm that follows
  const void* scale[2];
  long long gate_stride[2], mod_stride[2];
  void* xm[2][8];          // every rank's modulated-activation buffers
  float eps;
};

int tp_p2p_barrier(void* const* flags, int P, int rank, cudaStream_t st);
int tp_p2p_reduce_ln_push(const TpReduceArgs& a, cudaStream_t st);
// copy n bf16 elements (n % 8 == 0) from a local buffer to P destinations (own and peers'): a small all-gather
int tp_p2p_push_rows(const void* src, void* const* dst, long long n, int P, cudaStream_t st);

}  // namespace qimg
