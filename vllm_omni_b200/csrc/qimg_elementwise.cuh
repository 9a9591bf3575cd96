// HBM-bound kernels of the Qwen-Image DiT hot path (sm_100a): 128-bit coalesced accesses,
// one row per warp with the row held in registers, fp32 math with the reference's bf16
// rounding points reproduced (rbf).  Grids are sized from the row count; every kernel
// touches each byte exactly once (algorithmic bytes in DESIGN.md).
#pragma once

#include "qimg_common.cuh"

namespace qimg {

constexpr int EW_MAX_CHUNKS = 16;  // 16 x 256 = 4096 elements per row max (D=3072, joint_dim=3584)

// y = bf16( bf16( bf16(LN(x)) * bf16(1 + scale) ) + shift )
// Reference: AdaLayerNorm.forward_native, layers/adalayernorm.py:94-102 (LayerNorm no affine,
// eps 1e-6, then `* (1 + scale) + shift`, 4 separate ATen kernels there).
// shift/scale are [*, D] slices of the modulation buffer, row -> batch = row / rows_per_batch.
__global__ void __launch_bounds__(128) ln_modulate_kernel(const bf16* __restrict__ x, const bf16* __restrict__ shift,
                                                          const bf16* __restrict__ scale, bf16* __restrict__ y, int rows,
                                                          int row_base, int D, int rows_per_batch, long long mod_stride,
                                                          float eps, const int* __restrict__ skip,
                                                          const int* __restrict__ mod_index, int index_batch) {
  if (skip && *skip) return;  // step cache: this forward reuses the cached residual (qimg_tea_decide)
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nch = (D + 255) >> 8;
  const bf16* xr = x + (size_t)row * D;
  uint4 v[EW_MAX_CHUNKS];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < EW_MAX_CHUNKS; ++i) {
    const int e = i * 256 + lane * 8;
    if (i < nch && e < D) {
      v[i] = ldg_nc_v4(xr + e);
      sum += bf16lo(v[i].x) + bf16hi(v[i].x) + bf16lo(v[i].y) + bf16hi(v[i].y) + bf16lo(v[i].z) + bf16hi(v[i].z) +
             bf16lo(v[i].w) + bf16hi(v[i].w);
    }
  }
  const float mean = warp_sum(sum) / (float)D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < EW_MAX_CHUNKS; ++i) {
    const int e = i * 256 + lane * 8;
    if (i < nch && e < D) {
      uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float a = bf16lo(w[k]) - mean, b = bf16hi(w[k]) - mean;
        sq += a * a + b * b;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
  const float nmr = -mean * rstd;
  int b = (row_base + row) / rows_per_batch;
  // per-token modulation select (AdaLayerNorm.preprocess with `index`, layers/adalayernorm.py:31-54): the modulation
  // buffer holds 2 * index_batch rows, tokens with index != 0 take the second half
  if (mod_index && mod_index[row_base + row] != 0) b += index_batch;
  const bf16* sh = shift + (size_t)b * mod_stride;
  const bf16* sc = scale + (size_t)b * mod_stride;
  bf16* yr = y + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < EW_MAX_CHUNKS; ++i) {
    const int e = i * 256 + lane * 8;
    if (i < nch && e < D) {
      uint4 shv = __ldg(reinterpret_cast<const uint4*>(sh + e));
      uint4 scv = __ldg(reinterpret_cast<const uint4*>(sc + e));
      uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
      uint32_t s1[4] = {shv.x, shv.y, shv.z, shv.w}, s2[4] = {scv.x, scv.y, scv.z, scv.w}, o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // bf16(LN(x)) -> * bf16(1 + scale) -> + shift, each op rounded to bf16 (packed HMUL2/HADD2.BF16)
        const uint32_t n = pack_bf16x2(fmaf(bf16lo(w[k]), rstd, nmr), fmaf(bf16hi(w[k]), rstd, nmr));
        o[k] = badd2(bmul2(n, badd2(0x3F803F80u, s2[k])), s1[k]);
      }
      stg_v4(yr + e, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}

// Same op for D = NCH * 256 known at compile time (D = 3072 -> NCH = 12): no per-chunk predicates, the row is
// unpacked once into packed fp32x2 registers and all three passes (mean, variance, normalise+modulate) run on
// FADD2/FFMA2 + packed bf16 ops: ~5 instructions per element instead of ~15 (the generic kernel above is
// issue-bound at 2.4 TB/s; profiles/r01_ln_modulate).
template <int NCH>
__global__ void __launch_bounds__(128) ln_modulate_fast_kernel(const bf16* __restrict__ x, const bf16* __restrict__ shift,
                                                               const bf16* __restrict__ scale, bf16* __restrict__ y,
                                                               int rows, int row_base, int rows_per_batch,
                                                               long long mod_stride, float eps,
                                                               const int* __restrict__ skip,
                                                               const int* __restrict__ mod_index, int index_batch) {
  if (skip && *skip) return;
  constexpr int D = NCH * 256;
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const bf16* xr = x + (size_t)row * D + lane * 8;
  uint64_t c[NCH * 4];  // the row as fp32 pairs
  {
    uint4 v[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) v[i] = ldg_nc_v4(xr + i * 256);
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) c[i * 4 + k] = ew_pack2(w[k] << 16, w[k] & 0xffff0000u);
    }
  }
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
  const float rstd = rsqrtf(warp_sum(ew_hsum2(ew_add2(q0, q1))) * (1.0f / (float)D) + eps);
  const uint64_t rstd2 = ew_splat2(rstd), zero2 = 0;
  int b = (row_base + row) / rows_per_batch;
  if (mod_index && mod_index[row_base + row] != 0) b += index_batch;  // per-token modulation select (see the generic kernel)
  const bf16* sh = shift + (size_t)b * mod_stride + lane * 8;
  const bf16* sc = scale + (size_t)b * mod_stride + lane * 8;
  bf16* yr = y + (size_t)row * D + lane * 8;
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
    stg_v4(yr + i * 256, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// y = bf16( bf16(x * rsqrt(mean(x^2) + eps)) * w )     (vLLM RMSNorm, used for txt_norm :758)
__global__ void __launch_bounds__(128) rms_norm_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                       bf16* __restrict__ y, int rows, int D, float eps) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int nch = (D + 255) >> 8;
  const bf16* xr = x + (size_t)row * D;
  uint4 v[EW_MAX_CHUNKS];
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < EW_MAX_CHUNKS; ++i) {
    const int e = i * 256 + lane * 8;
    if (i < nch && e < D) {
      v[i] = ldg_nc_v4(xr + e);
      uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
      for (int k = 0; k < 4; ++k) sq += bf16lo(u[k]) * bf16lo(u[k]) + bf16hi(u[k]) * bf16hi(u[k]);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
  bf16* yr = y + (size_t)row * D;
#pragma unroll
  for (int i = 0; i < EW_MAX_CHUNKS; ++i) {
    const int e = i * 256 + lane * 8;
    if (i < nch && e < D) {
      uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + e));
      uint32_t u[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, ww[4] = {wv.x, wv.y, wv.z, wv.w}, o[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o[k] = pack_bf16x2(rbf(rbf(bf16lo(u[k]) * rstd) * bf16lo(ww[k])), rbf(rbf(bf16hi(u[k]) * rstd) * bf16hi(ww[k])));
      stg_v4(yr + e, make_uint4(o[0], o[1], o[2], o[3]));
    }
  }
}

// x = bf16( x + bf16(gate * y) )    standalone form of qwen_image_transformer.py:586-587,592,597
// (the hot path fuses this into the GEMM epilogue; this kernel backs the layer-level plug-in API)
__global__ void __launch_bounds__(256) gate_residual_kernel(bf16* __restrict__ x, const bf16* __restrict__ y,
                                                            const bf16* __restrict__ gate, long long n_vec, int D,
                                                            int rows_per_batch, long long gate_stride) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int dv = D >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const long long row = i / dv;
    const int col = (int)(i - row * dv) << 3;
    const long long b = row / rows_per_batch;
    uint4 xv = ldg_v4(x + i * 8), yv = ldg_nc_v4(y + i * 8);
    uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate + b * gate_stride + col));
    uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, yw[4] = {yv.x, yv.y, yv.z, yv.w}, gw[4] = {gv.x, gv.y, gv.z, gv.w}, o[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16x2(rbf(bf16lo(xw[k]) + rbf(bf16lo(gw[k]) * bf16lo(yw[k]))),
                         rbf(bf16hi(xw[k]) + rbf(bf16hi(gw[k]) * bf16hi(yw[k]))));
    stg_v4(x + i * 8, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// Tensor-parallel form: y is the all-reduced partial sum of a row-parallel linear (no bias yet):
//   x = bf16( x + bf16( gate * bf16(y + bias) ) )
__global__ void __launch_bounds__(256) gate_residual_bias_kernel(bf16* __restrict__ x, const bf16* __restrict__ y,
                                                                 const bf16* __restrict__ bias, const bf16* __restrict__ gate,
                                                                 long long n_vec, int D, int rows_per_batch,
                                                                 long long gate_stride) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const int dv = D >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const long long row = i / dv;
    const int col = (int)(i - row * dv) << 3;
    const long long b = row / rows_per_batch;
    const uint4 xv = ldg_v4(x + i * 8), yv = ldg_nc_v4(y + i * 8);
    const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate + b * gate_stride + col));
    const uint4 bv = __ldg(reinterpret_cast<const uint4*>(bias + col));
    uint4 o;
    o.x = badd2(xv.x, bmul2(gv.x, badd2(yv.x, bv.x)));
    o.y = badd2(xv.y, bmul2(gv.y, badd2(yv.y, bv.y)));
    o.z = badd2(xv.z, bmul2(gv.z, badd2(yv.z, bv.z)));
    o.w = badd2(xv.w, bmul2(gv.w, badd2(yv.w, bv.w)));
    stg_v4(x + i * 8, o);
  }
}

// ---- step-cache (TeaCache) helpers: reference cache/teacache/hook.py:124-160,196-206 ------------------------------
// sums[0] += sum |bf16(a - b)|, sums[1] += sum |b|  (fp32): the two means of
//   rel = (mod - prev).abs().mean() / (prev.abs().mean() + 1e-8)      (hook.py:198-203; the subtraction is a bf16 op)
// Each byte is read once; one atomicAdd pair per block.
__global__ void __launch_bounds__(256) rel_l1_sums_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b,
                                                          long long n_vec, float* __restrict__ sums) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  float sd = 0.f, sb = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 av = ldg_nc_v4(a + i * 8), bv = ldg_nc_v4(b + i * 8);
    const uint32_t aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t d = bsub2(aw[k], bw[k]);
      sd += fabsf(bf16lo(d)) + fabsf(bf16hi(d));
      sb += fabsf(bf16lo(bw[k])) + fabsf(bf16hi(bw[k]));
    }
  }
  sd = warp_sum(sd);
  sb = warp_sum(sb);
  __shared__ float red[2][8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = sd;
    red[1][warp] = sb;
  }
  __syncthreads();
  if (warp == 0) {
    sd = lane < 8 ? red[0][lane] : 0.f;
    sb = lane < 8 ? red[1][lane] : 0.f;
    sd = warp_sum(sd);
    sb = warp_sum(sb);
    if (lane == 0) {
      atomicAdd(sums, sd);
      atomicAdd(sums + 1, sb);
    }
  }
}

// out = bf16(a - b)  (cached residual, hook.py:152)   /   x = bf16(x + r)  (residual reuse, hook.py:131)
__global__ void __launch_bounds__(256) bf16_sub_kernel(bf16* __restrict__ out, const bf16* __restrict__ a,
                                                       const bf16* __restrict__ b, long long n_vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 av = ldg_nc_v4(a + i * 8), bv = ldg_nc_v4(b + i * 8);
    uint4 o;
    o.x = bsub2(av.x, bv.x);
    o.y = bsub2(av.y, bv.y);
    o.z = bsub2(av.z, bv.z);
    o.w = bsub2(av.w, bv.w);
    stg_v4(out + i * 8, o);
  }
}
__global__ void __launch_bounds__(256) bf16_add_inplace_kernel(bf16* __restrict__ x, const bf16* __restrict__ r, long long n_vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 xv = ldg_v4(x + i * 8), rv = ldg_nc_v4(r + i * 8);
    uint4 o;
    o.x = badd2(xv.x, rv.x);
    o.y = badd2(xv.y, rv.y);
    o.z = badd2(xv.z, rv.z);
    o.w = badd2(xv.w, rv.w);
    stg_v4(x + i * 8, o);
  }
}

// ---- TeaCache decision on the DEVICE (reference cache/teacache/hook.py:170-217 takes it on the host after `.cpu().item()`) ----
// One thread reproduces the host arithmetic bit for bit: the two means in fp32 rounded to bf16, the bf16 `+ 1e-8` and
// division, the degree-4 rescale polynomial (numpy.poly1d = Horner) and the accumulation in fp64.  flag = 1: reuse the
// cached residual (the engine's BLOCKS stage is launched predicated on it and exits at once), 0: compute.
struct TeaCoef {
  double c[5];  // highest power first
};
__global__ void tea_decide_kernel(const float* __restrict__ sums, double n, TeaCoef coef, double thresh, double* __restrict__ accum,
                                  int* __restrict__ flag, float* __restrict__ hist, int hist_idx, int force) {
  if (threadIdx.x || blockIdx.x) return;
  float rel = __int_as_float(0x7fc00000);  // NaN: no distance on a forced step (as the host hook records)
  int reuse = 0;
  if (force == 1) {  // first forward of a branch: reset the accumulator and compute (hook.py:183-186)
    *accum = 0.0;
  } else if (force == 0) {
    const float nf = (float)n;
    const float num = rbf(__fdiv_rn(sums[0], nf));
    const float den = rbf(__fdiv_rn(sums[1], nf));
    rel = rbf(__fdiv_rn(num, rbf(den + 1e-8f)));
    const double x = (double)rel;
    double y = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) y = y * x + coef.c[i];
    const double a = *accum + fabs(y);
    if (a < thresh) {
      *accum = a;
      reuse = 1;
    } else {
      *accum = 0.0;
    }
  }  // force == 2: no previous modulated input yet -> compute, accumulator untouched
  *flag = reuse;
  if (hist) {
    hist[2 * hist_idx] = (float)reuse;
    hist[2 * hist_idx + 1] = rel;
  }
}

// reuse (flag = 1): x += resid (hook.py:131-133);  compute (flag = 0): resid = x - ori (hook.py:152-154).  One pass either way.
__global__ void __launch_bounds__(256) tea_residual_kernel(bf16* __restrict__ x, const bf16* __restrict__ ori, bf16* __restrict__ resid,
                                                           long long n_vec, const int* __restrict__ flag) {
  const bool reuse = *flag != 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
    const uint4 xv = ldg_v4(x + i * 8);
    uint4 o;
    if (reuse) {
      const uint4 rv = ldg_nc_v4(resid + i * 8);
      o.x = badd2(xv.x, rv.x); o.y = badd2(xv.y, rv.y); o.z = badd2(xv.z, rv.z); o.w = badd2(xv.w, rv.w);
      stg_v4(x + i * 8, o);
    } else {
      const uint4 ov = ldg_nc_v4(ori + i * 8);
      o.x = bsub2(xv.x, ov.x); o.y = bsub2(xv.y, ov.y); o.z = bsub2(xv.z, ov.z); o.w = bsub2(xv.w, ov.w);
      stg_v4(resid + i * 8, o);
    }
  }
}

// Small-M linear ("GEMV"): y[m, n] = bf16( sum_k act(x[m,k]) * W[n,k] + bias[n] ), M <= 8 per pass.
// HBM-bound on W (read exactly once).  Used for the timestep MLP (qwen_image_transformer.py:50-62),
// all 2*L modulation projections img_mod/txt_mod (:552-557, batched into ONE launch over the
// concatenated [L*2*6D, D] weight) and norm_out.linear (:797).  act = SiLU on the input when ACT_SILU.
template <int MAXM>
__global__ void __launch_bounds__(256) linear_small_m_kernel(const bf16* __restrict__ x, const bf16* __restrict__ W,
                                                             const bf16* __restrict__ bias, bf16* __restrict__ y, int M,
                                                             long long N, int K, long long ldy, int act_silu) {
  extern __shared__ uint8_t smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);  // [M][K], activation applied, bf16-rounded like the reference
  for (int i = threadIdx.x; i < M * K; i += blockDim.x) {
    float v = __bfloat162float(x[i]);
    if (act_silu) v = rbf(silu_f(v));
    xs[i] = __float2bfloat16_rn(v);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
  const int kv = K >> 3;  // 16-byte vectors per row
  for (long long n = (long long)blockIdx.x * (blockDim.x >> 5) + warp; n < N; n += warps_total) {
    const bf16* wr = W + n * K;
    float acc[MAXM];
#pragma unroll
    for (int m = 0; m < MAXM; ++m) acc[m] = 0.f;
    for (int c = lane; c < kv; c += 32) {
      uint4 wv = ldg_nc_v4(wr + c * 8);
      uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
      for (int m = 0; m < MAXM; ++m) {
        if (m < M) {
          uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)m * K + c * 8);
          uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            acc[m] = fmaf(bf16lo(ww[k]), bf16lo(xw[k]), acc[m]);
            acc[m] = fmaf(bf16hi(ww[k]), bf16hi(xw[k]), acc[m]);
          }
        }
      }
    }
    const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
#pragma unroll
    for (int m = 0; m < MAXM; ++m) {
      if (m < M) {
        float s = warp_sum(acc[m]);
        if (lane == 0) y[(size_t)m * ldy + n] = __float2bfloat16_rn(s + bv);
      }
    }
  }
}

// Timesteps(256, flip_sin_to_cos=True, downscale_freq_shift=0, scale=1000) -> bf16 [B,256] = [cos | sin]
// (qwen_image_transformer.py:44,51-52; restated in-tree at pipeline_qwen_image.py:135-184)
__global__ void timestep_sinusoid_kernel(const bf16* __restrict__ t, bf16* __restrict__ out, int B) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 256) return;
  const int b = i >> 8, j = i & 255;
  const int f = j & 127;
  const float freq = expf(-9.210340371976184f * (float)f / 128.0f);
  const float arg = 1000.0f * (__bfloat162float(t[b]) * freq);
  out[i] = __float2bfloat16_rn(j < 128 ? cosf(arg) : sinf(arg));
}

// Fused true-CFG combine + norm rescale + flow-match Euler update, one pass over the latents.
//   comb = neg + s*(pos-neg); noise = comb * (||pos|| / ||comb||)   pipeline_qwen_image.py:580-583
//   x    = bf16( float(x) + bf16(bf16(dt) * noise) )                FlowMatchEulerDiscreteScheduler.step (:585)
//          (dt is a 0-dim fp32 tensor: under torch type promotion `dt * model_output` is computed in
//           model_output's dtype, i.e. dt is first rounded to bf16 — verified against torch on CPU)
// Rows have C = 64 channels (128 B): 8 lanes x 16 B per row, 4 rows per warp instruction.
// neg == nullptr -> no CFG (noise = pos).
__global__ void __launch_bounds__(256) cfg_euler_step_kernel(const bf16* __restrict__ pos, const bf16* __restrict__ neg,
                                                             bf16* __restrict__ x, long long rows, float cfg_scale,
                                                             float dt, const float* __restrict__ sigma_pair, int dt_fp32) {
  // sigma_pair != nullptr: (sigma_i, sigma_{i+1}) live in device memory, so a CUDA graph of the step can be replayed
  // for every timestep (the fp32 subtraction is the same one the host path does)
  if (sigma_pair) dt = __ldg(sigma_pair + 1) - __ldg(sigma_pair);
  const long long vec = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte vector per thread
  const long long row = vec >> 3;
  const bool active = row < rows;
  uint32_t pw[4] = {0, 0, 0, 0}, nw[4] = {0, 0, 0, 0}, xw[4] = {0, 0, 0, 0};
  if (active) {
    uint4 pv = ldg_nc_v4(pos + vec * 8);
    pw[0] = pv.x; pw[1] = pv.y; pw[2] = pv.z; pw[3] = pv.w;
    uint4 xv = ldg_v4(x + vec * 8);
    xw[0] = xv.x; xw[1] = xv.y; xw[2] = xv.z; xw[3] = xv.w;
    if (neg) {
      uint4 nv = ldg_nc_v4(neg + vec * 8);
      nw[0] = nv.x; nw[1] = nv.y; nw[2] = nv.z; nw[3] = nv.w;
    }
  }
  float noise[8];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    noise[2 * k] = bf16lo(pw[k]);
    noise[2 * k + 1] = bf16hi(pw[k]);
  }
  if (neg) {  // uniform branch
    float comb[8], pp = 0.f, cc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float n0 = bf16lo(nw[k]), n1 = bf16hi(nw[k]);
      comb[2 * k] = rbf(n0 + rbf(cfg_scale * rbf(noise[2 * k] - n0)));
      comb[2 * k + 1] = rbf(n1 + rbf(cfg_scale * rbf(noise[2 * k + 1] - n1)));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      pp += noise[k] * noise[k];
      cc += comb[k] * comb[k];
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
      cc += __shfl_xor_sync(0xffffffffu, cc, o);
    }
    const float ratio = rbf(rbf(sqrtf(pp)) / rbf(sqrtf(cc)));
#pragma unroll
    for (int k = 0; k < 8; ++k) noise[k] = rbf(comb[k] * ratio);
  }
  if (active) {
    uint32_t o[4];
    // the scheduler's `dt * model_output` multiplies a 0-dim fp32 tensor with a bf16 tensor: type promotion keeps bf16,
    // and a 0-dim DEVICE tensor (diffusers keeps sigmas on the device) is cast to bf16 inside the kernel before the
    // multiply (dt_fp32 == 0, the default); a 0-dim CPU tensor would be used at fp32 instead (dt_fp32 != 0)
    const float dtb = dt_fp32 ? dt : rbf(dt);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      o[k] = pack_bf16x2(bf16lo(xw[k]) + rbf(dtb * noise[2 * k]), bf16hi(xw[k]) + rbf(dtb * noise[2 * k + 1]));
    stg_v4(x + vec * 8, make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// out[r, :] = src[(index[r] != 0 ? index_batch : 0) + r / rows_per_batch, :]  — the per-token gate of AdaLayerNorm.preprocess
// (`torch.where(index == 0, gate_0, gate_1)`, layers/adalayernorm.py:47-49); D % 8 == 0, one uint4 per thread
__global__ void select_rows_kernel(const bf16* __restrict__ src, long long src_stride, const int* __restrict__ index,
                                   bf16* __restrict__ out, long long n_vec, int D, int rows_per_batch, int index_batch) {
  const int vpr = D / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / vpr;
    const int c = (int)(i - r * vpr);
    const long long b = r / rows_per_batch + (index[r] != 0 ? index_batch : 0);
    reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(src + b * src_stride) + c);
  }
}

}  // namespace qimg
