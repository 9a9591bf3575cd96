// Whole-model DiT forward (QwenImageTransformer2DModel.forward, reference
// qwen_image_transformer.py:692-802) as a native launch sequence over the sm_100a kernels:
// per block 2+2 LN-modulate, 4 grouped tcgen05 GEMMs and 1 joint attention = 9 launches
// (the reference issues ~55-60 eager kernels per block).  No host synchronisation, no
// allocation: the caller provides the workspace and the stream, so the whole forward is
// CUDA-graph capturable.
#include "../../include/qimg_b200.h"

#include <nvtx3/nvToolsExt.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "qimg_host.cuh"
#include "qimg_tp.h"

using namespace qimg;

struct qimg_engine {
  qimg_dims dims;
  qimg_global_weights g;
  std::vector<qimg_block_weights> blocks;
  int tp_size = 1;
  qimg_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  // peer-memory TP (qimg_engine_set_tp_p2p): every rank's workspace and barrier flags, mapped into this process
  bool p2p = false;
  // sequence parallelism (qimg_engine_set_sp_p2p): full weights, own rows through the linears, local heads through attention
  int sp_size = 1;
  int sp_rank = 0;
  const int* blocks_predicate = nullptr;  // qimg_engine_set_blocks_predicate
  bool p2p_ready = false;  // peer pointers registered (qimg_engine_set_tp_p2p with non-NULL arrays)
  int tp_rank = 0;
  void* peer_ws[8] = {};
  void* peer_flags[8] = {};
};

namespace {

// NVTX ranges per launch family (SURVEY §5 tracing row): off by default (qimg_set_nvtx / env QIMG_NVTX=1); header-only NVTX3,
// a no-op unless a profiler injects its library.
int g_nvtx = -1;
inline bool nvtx_on() {
  if (g_nvtx < 0) {
    const char* e = getenv("QIMG_NVTX");
    g_nvtx = (e && atoi(e) != 0) ? 1 : 0;
  }
  return g_nvtx == 1;
}
struct NvtxRange {
  bool on;
  explicit NvtxRange(const char* name) : on(nvtx_on()) {
    if (on) nvtxRangePushA(name);
  }
  ~NvtxRange() {
    if (on) nvtxRangePop();
  }
};
#define QIMG_RANGE_CAT2(a, b) a##b
#define QIMG_RANGE_CAT(a, b) QIMG_RANGE_CAT2(a, b)
#define QIMG_RANGE(name) NvtxRange QIMG_RANGE_CAT(_nvtx_range_, __LINE__)(name)

inline size_t align_up(size_t x, size_t a = 1024) { return (x + a - 1) / a * a; }

struct WsLayout {
  size_t x_img, x_txt, xm_img, xm_txt, q, k, v, at_img, at_txt, h_img, h_txt, txt_normed, tsin, t1, temb, mod_all, emb_out,
      part, zero_bias, recv, out_own, out_all, total;
  int own_img, own_txt;  // peer-memory TP: largest owner slice of the image / text rows
};

// tp > 1: head-sharded q/k/v/attention-output and FF-sharded MLP hidden buffers are 1/tp of the full size
// sp > 1 (sequence parallel): same head split for q/k/v; attention output and MLP hidden hold OWN rows at full width
WsLayout ws_layout(const qimg_dims& d, int B, int S_img, int T, int n_t_max, int tp = 1, bool p2p = false, int sp = 1) {
  if (sp > 1) tp = sp;
  const size_t D = (size_t)d.num_heads * d.head_dim, FF = 4 * D, S = (size_t)S_img + T;
  const size_t Mi = (size_t)B * S_img, Mt = (size_t)B * T;
  const size_t Hl = (size_t)d.num_heads / tp, Dl = D / tp, FFl = FF / tp;
  WsLayout w;
  size_t off = 0;
  auto take = [&](size_t elems) {
    size_t o = off;
    off = align_up(off + elems * 2);
    return o;
  };
  w.x_img = take(Mi * D);
  w.x_txt = take(Mt * D);
  w.xm_img = take(Mi * D);
  w.xm_txt = take(Mt * D);
  w.q = take((size_t)B * Hl * S * 128);
  w.k = take((size_t)B * Hl * S * 128);
  w.v = take((size_t)B * Hl * S * 128);
  const size_t oi = (Mi + tp - 1) / tp, ot = (Mt + tp - 1) / tp;  // largest owner slice (sequence parallel)
  w.at_img = take(sp > 1 ? oi * D : Mi * Dl);
  w.at_txt = take(sp > 1 ? ot * D : Mt * Dl);
  w.h_img = take(sp > 1 ? oi * FF : Mi * FFl);
  w.h_txt = take(sp > 1 ? ot * FF : Mt * FFl);
  w.txt_normed = take(Mt * d.joint_dim);
  w.tsin = take((size_t)n_t_max * 256);
  w.t1 = take((size_t)n_t_max * D);
  w.temb = take((size_t)n_t_max * D);
  w.mod_all = take((size_t)n_t_max * d.num_layers * 12 * D);
  w.emb_out = take((size_t)n_t_max * 2 * D);
  const bool nccl_tp = tp > 1 && !p2p && sp <= 1;
  w.part = take(nccl_tp ? (Mi + Mt) * D : 0);  // NCCL mode: [img rows | txt rows] x D bf16 partial sums of the row-parallel linears
  w.zero_bias = take(nccl_tp ? D : 0);
  // peer-memory mode: fp32 partial sums of MY rows from every source rank, [tp][own_img + own_txt][D]
  w.own_img = tp > 1 ? (int)((Mi + tp - 1) / tp) : 0;
  w.own_txt = tp > 1 ? (int)((Mt + tp - 1) / tp) : 0;
  w.recv = take((tp > 1 && p2p && sp <= 1) ? (size_t)tp * (w.own_img + w.own_txt) * D * 2 : 0);
  w.out_own = take(sp > 1 ? oi * d.out_dim : 0);   // sequence parallel: proj_out of my rows ...
  w.out_all = take(sp > 1 ? Mi * d.out_dim : 0);   // ... and every rank's rows, pushed by their owners
  w.total = off;
  return w;
}

}  // namespace

extern "C" {

void qimg_set_nvtx(int on) { g_nvtx = on != 0; }

int qimg_engine_set_blocks_predicate(qimg_engine* e, const int* skip_flag) {
  if (!e) return fail("qimg_engine_set_blocks_predicate: null engine");
  e->blocks_predicate = skip_flag;
  return 0;
}

int qimg_engine_create(const qimg_dims* dims, const qimg_global_weights* g, const qimg_block_weights* blocks,
                       qimg_engine** out) {
  if (!dims || !g || !blocks || !out) return fail("qimg_engine_create: null argument");
  if (dims->head_dim != 128) return fail("qimg_engine_create: head_dim must be 128");
  if (dims->num_layers <= 0 || dims->num_heads <= 0) return fail("qimg_engine_create: bad dims");
  if (dims->in_channels % 8 || dims->joint_dim % 8 || dims->out_dim % 8) return fail("qimg_engine_create: channel dims must be multiples of 8");
  qimg_engine* e = new (std::nothrow) qimg_engine();
  if (!e) return fail("qimg_engine_create: out of memory");
  e->dims = *dims;
  e->g = *g;
  e->blocks.assign(blocks, blocks + dims->num_layers);
  *out = e;
  return 0;
}

void qimg_engine_destroy(qimg_engine* e) { delete e; }

int qimg_engine_set_tp(qimg_engine* e, int tp_size, qimg_allreduce_fn allreduce, void* user) {
  if (!e) return fail("qimg_engine_set_tp: null engine");
  if (tp_size < 1 || e->dims.num_heads % tp_size) return fail("qimg_engine_set_tp: tp_size must divide num_heads");
  if (tp_size > 1 && !allreduce) return fail("qimg_engine_set_tp: an all-reduce callback is required for tp_size > 1");
  e->tp_size = tp_size;
  e->allreduce = allreduce;
  e->allreduce_user = user;
  e->p2p = false;
  return 0;
}

int qimg_engine_set_tp_p2p(qimg_engine* e, int tp_size, int tp_rank, void* const* peer_workspaces, void* const* peer_flags) {
  if (!e) return fail("qimg_engine_set_tp_p2p: null engine");
  if (tp_size != 2 && tp_size != 4 && tp_size != 8) return fail("qimg_engine_set_tp_p2p: tp_size must be 2, 4 or 8");
  if (e->dims.num_heads % tp_size) return fail("qimg_engine_set_tp_p2p: tp_size must divide num_heads");
  if (tp_rank < 0 || tp_rank >= tp_size) return fail("qimg_engine_set_tp_p2p: bad rank");
  if ((peer_workspaces == nullptr) != (peer_flags == nullptr)) return fail("qimg_engine_set_tp_p2p: pass both pointer arrays or neither");
  e->p2p_ready = false;
  if (peer_workspaces) {
    for (int p = 0; p < tp_size; ++p) {
      if (!peer_workspaces[p] || !peer_flags[p]) return fail("qimg_engine_set_tp_p2p: null peer pointer");
      e->peer_ws[p] = peer_workspaces[p];
      e->peer_flags[p] = peer_flags[p];
    }
    e->p2p_ready = true;
  }
  e->tp_size = tp_size;
  e->tp_rank = tp_rank;
  e->p2p = true;
  return 0;
}

int qimg_engine_set_sp_p2p(qimg_engine* e, int sp_size, int sp_rank, void* const* peer_workspaces, void* const* peer_flags) {
  if (!e) return fail("qimg_engine_set_sp_p2p: null engine");
  if (sp_size != 2 && sp_size != 4 && sp_size != 8) return fail("qimg_engine_set_sp_p2p: sp_size must be 2, 4 or 8");
  if (e->dims.num_heads % sp_size) return fail("qimg_engine_set_sp_p2p: sp_size must divide num_heads");
  if (sp_rank < 0 || sp_rank >= sp_size) return fail("qimg_engine_set_sp_p2p: bad rank");
  if (e->tp_size > 1) return fail("qimg_engine_set_sp_p2p: the engine is already tensor parallel");
  if ((peer_workspaces == nullptr) != (peer_flags == nullptr)) return fail("qimg_engine_set_sp_p2p: pass both pointer arrays or neither");
  e->p2p_ready = false;
  if (peer_workspaces) {
    for (int p = 0; p < sp_size; ++p) {
      if (!peer_workspaces[p] || !peer_flags[p]) return fail("qimg_engine_set_sp_p2p: null peer pointer");
      e->peer_ws[p] = peer_workspaces[p];
      e->peer_flags[p] = peer_flags[p];
    }
    e->p2p_ready = true;
  }
  e->sp_size = sp_size;
  e->sp_rank = sp_rank;
  e->tp_rank = sp_rank;  // barrier / error plumbing shared with the TP mode
  return 0;
}

int qimg_engine_p2p_error(qimg_engine* e, int* out) {
  if (!e || !(e->p2p || e->sp_size > 1) || !e->p2p_ready || !out) return fail("qimg_engine_p2p_error: engine is not in a peer-memory mode");
  QIMG_CUDA_CHECK(cudaMemcpy(out, (char*)e->peer_flags[e->tp_rank] + 64, sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}

size_t qimg_engine_workspace_bytes(const qimg_engine* e, int B, int S_img, int T) {
  return ws_layout(e->dims, B, S_img, T, B, e->tp_size, e->p2p, e->sp_size).total;
}
size_t qimg_engine_ws_offset_img(const qimg_engine* e, int B, int S_img, int T) { return ws_layout(e->dims, B, S_img, T, B, e->tp_size, e->p2p, e->sp_size).x_img; }
size_t qimg_engine_ws_offset_txt(const qimg_engine* e, int B, int S_img, int T) { return ws_layout(e->dims, B, S_img, T, B, e->tp_size, e->p2p, e->sp_size).x_txt; }

#define QIMG_TRY(expr)      \
  do {                      \
    int _rc = (expr);       \
    if (_rc) return _rc;    \
  } while (0)

int qimg_engine_forward(qimg_engine* e, const void* hidden, const void* enc, const void* timestep, int n_t,
                        const void* img_cos, const void* img_sin, const void* txt_cos, const void* txt_sin, int B,
                        int S_img, int T, void* out, void* workspace, size_t workspace_bytes, qimg_stream_t st) {
  return qimg_engine_forward_stages(e, QIMG_STAGE_ALL, hidden, enc, timestep, n_t, img_cos, img_sin, txt_cos, txt_sin, B, S_img, T,
                                    out, workspace, workspace_bytes, st);
}

size_t qimg_engine_ws_offset_mod(const qimg_engine* e, int B, int S_img, int T) { return ws_layout(e->dims, B, S_img, T, B, e->tp_size, e->p2p, e->sp_size).xm_img; }

int qimg_engine_forward_stages(qimg_engine* e, int stages, const void* hidden, const void* enc, const void* timestep, int n_t,
                               const void* img_cos, const void* img_sin, const void* txt_cos, const void* txt_sin, int B,
                               int S_img, int T, void* out, void* workspace, size_t workspace_bytes, qimg_stream_t st) {
  if (!e) return fail("qimg_engine_forward: null engine");
  if (stages <= 0 || stages > QIMG_STAGE_ALL) return fail("qimg_engine_forward_stages: bad stage mask");
  if (B <= 0 || S_img <= 0 || T <= 0) return fail("qimg_engine_forward: bad shape");
  if (n_t != 1 && n_t != B) return fail("qimg_engine_forward: n_t must be 1 or B");
  const qimg_dims& d = e->dims;
  const int tp = e->tp_size;
  const WsLayout w = ws_layout(d, B, S_img, T, B, tp, e->p2p, e->sp_size);
  if (workspace_bytes < w.total) return fail("qimg_engine_forward: workspace too small");
  if (reinterpret_cast<uintptr_t>(workspace) & 1023) return fail("qimg_engine_forward: workspace must be 1024-byte aligned");
  char* ws = static_cast<char*>(workspace);
  const int H = d.num_heads, D = H * 128, FF = 4 * D, L = d.num_layers, S = S_img + T;
  const int Hl = H / tp, Dl = D / tp, FFl = FF / tp;  // local heads / widths under tensor parallelism
  void *part = ws + w.part, *zero_bias = ws + w.zero_bias;
  char* part_txt = (char*)part + (size_t)B * S_img * D * 2;
  if (tp > 1 && !e->p2p) QIMG_CUDA_CHECK(cudaMemsetAsync(zero_bias, 0, (size_t)D * 2, (cudaStream_t)st));
  const int sp = e->sp_size;
  if (sp > 1) {
    if (!e->p2p_ready) return fail("qimg_engine_forward: sequence parallelism declared but no peer workspaces registered");
    if (workspace != e->peer_ws[e->sp_rank]) return fail("qimg_engine_forward: sequence parallelism needs the registered workspace");
    if (stages != QIMG_STAGE_ALL) return fail("qimg_engine_forward_stages: staged forwards are not available under sequence parallelism");
  }
  if (e->p2p) {
    if (!e->p2p_ready) return fail("qimg_engine_forward: peer-memory TP declared but no peer workspaces registered");
    if (workspace != e->peer_ws[e->tp_rank]) return fail("qimg_engine_forward: peer-memory TP needs the registered workspace");
    if (stages != QIMG_STAGE_ALL) return fail("qimg_engine_forward_stages: staged forwards are not available in peer-memory TP mode");
  }
  // NCCL mode: bf16 partial sums -> caller's all-reduce -> x += gate * (sum + bias)   (comparison baseline)
  auto tp_reduce = [&](const void* b_img, const void* g_img, const void* b_txt, const void* g_txt, long long gstride) -> int {
    if (!e->allreduce) return fail("TP: no all-reduce registered");
    if (e->allreduce(part, (long long)(B * S_img + B * T) * D, e->allreduce_user, st)) return fail("TP all-reduce callback failed");
    QIMG_TRY(qimg_gate_residual_bias(ws + w.x_img, part, b_img, g_img, B * S_img, D, S_img, gstride, st));
    QIMG_TRY(qimg_gate_residual_bias(ws + w.x_txt, part_txt, b_txt, g_txt, B * T, D, T, gstride, st));
    return 0;
  };
  // peer-memory mode (csrc/qimg_tp_p2p.cu): the row-parallel GEMM pushed fp32 partial sums to the row owners; each rank
  // now reduces ITS rows, applies bias + gate + residual, runs the NEXT AdaLayerNorm on the row in registers and stores
  // the modulated row into every rank's xm buffer; two cross-GPU barriers order pushes / reads
  auto tp_reduce_ln = [&](const void* b_img, const void* g_img, const void* b_txt, const void* g_txt, long long gstride,
                          const void* sh_img, const void* sc_img, const void* sh_txt, const void* sc_txt,
                          long long mstride) -> int {
    TpReduceArgs a;
    memset(&a, 0, sizeof a);
    a.P = tp; a.rank = e->tp_rank; a.D = D; a.eps = d.eps;
    a.recv_local = ws + w.recv;
    a.recv_rows = w.own_img + w.own_txt;
    a.row_off[0] = 0; a.row_off[1] = w.own_img;
    a.x[0] = ws + w.x_img; a.x[1] = ws + w.x_txt;
    a.rows[0] = B * S_img; a.rows[1] = B * T;
    a.rows_per_batch[0] = S_img; a.rows_per_batch[1] = T;
    a.bias[0] = b_img; a.bias[1] = b_txt;
    a.gate[0] = g_img; a.gate[1] = g_txt;
    a.shift[0] = sh_img; a.shift[1] = sh_txt;
    a.scale[0] = sc_img; a.scale[1] = sc_txt;
    a.gate_stride[0] = a.gate_stride[1] = gstride;
    a.mod_stride[0] = a.mod_stride[1] = mstride;
    for (int p = 0; p < tp; ++p) {
      a.xm[0][p] = (char*)e->peer_ws[p] + w.xm_img;
      a.xm[1][p] = (char*)e->peer_ws[p] + w.xm_txt;
    }
    QIMG_TRY(tp_p2p_barrier(e->peer_flags, tp, e->tp_rank, (cudaStream_t)st));   // every rank's pushes have landed
    QIMG_TRY(tp_p2p_reduce_ln_push(a, (cudaStream_t)st));
    QIMG_TRY(tp_p2p_barrier(e->peer_flags, tp, e->tp_rank, (cudaStream_t)st));   // every rank's xm rows have landed
    return 0;
  };
  auto partial_problem = [&](qimg_gemm_problem& p, int row_off) {
    p.tp_size = tp; p.tp_rank = e->tp_rank; p.tp_recv_rows = w.own_img + w.own_txt; p.tp_recv_row_off = row_off;
    for (int r = 0; r < tp; ++r) p.tp_recv[r] = (char*)e->peer_ws[r] + w.recv;
  };
  const int Mi = B * S_img, Mt = B * T;
  void *x_img = ws + w.x_img, *x_txt = ws + w.x_txt, *xm_img = ws + w.xm_img, *xm_txt = ws + w.xm_txt;
  void *q = ws + w.q, *k = ws + w.k, *v = ws + w.v, *at_img = ws + w.at_img, *at_txt = ws + w.at_txt;
  void *h_img = ws + w.h_img, *h_txt = ws + w.h_txt, *txt_normed = ws + w.txt_normed;
  void *tsin = ws + w.tsin, *t1 = ws + w.t1, *temb = ws + w.temb, *emb_out = ws + w.emb_out;
  char* mod_all = ws + w.mod_all;
  const long long mod_ld = (long long)L * 12 * D;               // elements per timestep row of mod_all
  const long long mod_stride = (n_t == 1) ? 0 : mod_ld;         // shared timestep -> one modulation row
  const long long emb_stride = (n_t == 1) ? 0 : 2LL * D;

  QIMG_RANGE("qimg.forward");
  // ---- prologue: temb, all modulations, img_in, txt_norm + txt_in -------------------------
  if (stages & QIMG_STAGE_PRE) {
  QIMG_RANGE("qimg.pre (temb, modulations, img_in, txt_in)");
  QIMG_TRY(qimg_timestep_sinusoid(timestep, tsin, n_t, st));
  QIMG_TRY(qimg_linear_small_m(tsin, e->g.t_lin1_w, e->g.t_lin1_b, t1, n_t, D, 256, D, 0, st));
  QIMG_TRY(qimg_linear_small_m(t1, e->g.t_lin2_w, e->g.t_lin2_b, temb, n_t, D, D, D, 1, st));
  if (e->g.mod_all_w) {
    QIMG_TRY(qimg_linear_small_m(temb, e->g.mod_all_w, e->g.mod_all_b, mod_all, n_t, (long long)L * 12 * D, D, mod_ld, 1, st));
  } else {
    for (int l = 0; l < L; ++l) {
      const qimg_block_weights& bw = e->blocks[l];
      QIMG_TRY(qimg_linear_small_m(temb, bw.img_mod_w, bw.img_mod_b, mod_all + ((size_t)l * 12 * D) * 2, n_t, 6LL * D, D, mod_ld, 1, st));
      QIMG_TRY(qimg_linear_small_m(temb, bw.txt_mod_w, bw.txt_mod_b, mod_all + ((size_t)l * 12 * D + 6 * D) * 2, n_t, 6LL * D, D, mod_ld, 1, st));
    }
  }
  QIMG_TRY(qimg_linear_small_m(temb, e->g.norm_out_w, e->g.norm_out_b, emb_out, n_t, 2LL * D, D, 2LL * D, 1, st));
  QIMG_TRY(qimg_rms_norm(enc, e->g.txt_norm_w, txt_normed, Mt, d.joint_dim, d.eps, st));
  {
    qimg_gemm_problem p[2];
    memset(p, 0, sizeof p);
    p[0].A = hidden; p[0].W = e->g.img_in_w; p[0].bias = e->g.img_in_b;
    p[0].M = Mi; p[0].N = D; p[0].K = d.in_channels; p[0].rows_per_batch = S_img; p[0].out = x_img; p[0].ldo = D;
    p[1].A = txt_normed; p[1].W = e->g.txt_in_w; p[1].bias = e->g.txt_in_b;
    p[1].M = Mt; p[1].N = D; p[1].K = d.joint_dim; p[1].rows_per_batch = T; p[1].out = x_txt; p[1].ldo = D;
    QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS, st));
  }
  if (!(stages & QIMG_STAGE_BLOCKS)) {
    // staged call (step caches): leave block 0's modulated image stream in the workspace — the quantity TeaCache
    // compares between steps (reference cache/teacache/extractors.py:206-209)
    const char* mi0 = mod_all;
    QIMG_TRY(qimg_ln_modulate(x_img, mi0, mi0 + (size_t)D * 2, xm_img, Mi, D, S_img, mod_stride, d.eps, st));
  }
  }  // QIMG_STAGE_PRE

  // ---- sequence parallel (Ulysses, fused): own rows through every linear with the FULL weights, local heads through
  //      attention; the two all-to-alls are peer stores of the QKV-GEMM and attention epilogues -----------------------------
  if (sp > 1) {
    const int r = e->sp_rank;
    auto own = [&](int rows, int* r0, int* n) {
      const int base = rows / sp, extra = rows % sp;
      *r0 = r * base + (r < extra ? r : extra);
      *n = base + (r < extra ? 1 : 0);
    };
    int i0, ni, t0, nt;
    own(Mi, &i0, &ni);
    own(Mt, &t0, &nt);
    const int Hs = H / sp;
    auto rows_at = [&](void* base, int row0, int width) { return (void*)((char*)base + (size_t)row0 * width * 2); };
    void *xi = rows_at(x_img, i0, D), *xt = rows_at(x_txt, t0, D), *xmi = rows_at(xm_img, i0, D), *xmt = rows_at(xm_txt, t0, D);
    const float sm_scale_sp = 1.0f / sqrtf(128.0f);
    for (int l = 0; l < L; ++l) {
      QIMG_RANGE("qimg.block (sequence parallel)");
      const qimg_block_weights& bw = e->blocks[l];
      const char* mi = mod_all + ((size_t)l * 12 * D) * 2;
      const char* mt = mi + (size_t)6 * D * 2;
      auto seg = [&](const char* base, int i) { return (const void*)(base + (size_t)i * D * 2); };
      QIMG_TRY(qimg_ln_modulate_rows(xi, seg(mi, 0), seg(mi, 1), xmi, ni, i0, D, S_img, mod_stride, d.eps, st));
      QIMG_TRY(qimg_ln_modulate_rows(xt, seg(mt, 0), seg(mt, 1), xmt, nt, t0, D, T, mod_stride, d.eps, st));
      {  // QKV of MY rows for ALL heads; each head's rows land in its owner's joint q / k / v (all-to-all #1)
        qimg_gemm_problem p[2];
        memset(p, 0, sizeof p);
        p[0].A = xmi; p[0].W = bw.to_qkv_w; p[0].bias = bw.to_qkv_b; p[0].M = ni; p[0].N = 3 * D; p[0].K = D;
        p[0].rows_per_batch = S_img; p[0].row_base = i0; p[0].norm_q_w = bw.norm_q; p[0].norm_k_w = bw.norm_k;
        p[0].rope_cos = img_cos; p[0].rope_sin = img_sin; p[0].S_joint = S; p[0].pos_off = T; p[0].H = H; p[0].eps = d.eps;
        p[0].sp_size = sp;
        for (int o = 0; o < sp; ++o) {
          p[0].sp_q[o] = (char*)e->peer_ws[o] + w.q; p[0].sp_k[o] = (char*)e->peer_ws[o] + w.k; p[0].sp_v[o] = (char*)e->peer_ws[o] + w.v;
        }
        p[1] = p[0];
        p[1].A = xmt; p[1].W = bw.add_kv_w; p[1].bias = bw.add_kv_b; p[1].M = nt; p[1].rows_per_batch = T; p[1].row_base = t0;
        p[1].norm_q_w = bw.norm_added_q; p[1].norm_k_w = bw.norm_added_k; p[1].rope_cos = txt_cos; p[1].rope_sin = txt_sin;
        p[1].pos_off = 0;
        if (nt > 0) QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_QKV, st));
        else QIMG_TRY(qimg_gemm(p, 1, QIMG_EPI_QKV, st));
      }
      QIMG_TRY(tp_p2p_barrier(e->peer_flags, sp, r, (cudaStream_t)st));  // every rank's q / k / v rows have landed
      {  // attention over my heads, all rows; output rows go to their owners' at buffers (all-to-all #2)
        qimg_fmha_sp fs;
        memset(&fs, 0, sizeof fs);
        fs.sp_size = sp; fs.sp_rank = r;
        for (int o = 0; o < sp; ++o) {
          fs.out_img[o] = (char*)e->peer_ws[o] + w.at_img;
          fs.out_txt[o] = (char*)e->peer_ws[o] + w.at_txt;
        }
        QIMG_TRY(qimg_fmha_joint_sp(q, k, v, B, Hs, S, T, sm_scale_sp, &fs, st));
      }
      QIMG_TRY(tp_p2p_barrier(e->peer_flags, sp, r, (cudaStream_t)st));  // every head's contribution to my rows has landed
      auto own_problem = [&](qimg_gemm_problem& p, const void* A, const void* W, const void* bias, int M, int N, int K, int rpb,
                             int row_base) {
        p.A = A; p.W = W; p.bias = bias; p.M = M; p.N = N; p.K = K; p.rows_per_batch = rpb; p.row_base = row_base;
      };
      {
        qimg_gemm_problem p[2];
        memset(p, 0, sizeof p);
        own_problem(p[0], at_img, bw.to_out_w, bw.to_out_b, ni, D, D, S_img, i0);
        p[0].out = xi; p[0].ldo = D; p[0].gate = seg(mi, 2); p[0].gate_stride = mod_stride;
        own_problem(p[1], at_txt, bw.to_add_out_w, bw.to_add_out_b, nt, D, D, T, t0);
        p[1].out = xt; p[1].ldo = D; p[1].gate = seg(mt, 2); p[1].gate_stride = mod_stride;
        QIMG_TRY(qimg_gemm(p, nt > 0 ? 2 : 1, QIMG_EPI_BIAS_GATE_RES, st));
      }
      QIMG_TRY(qimg_ln_modulate_rows(xi, seg(mi, 3), seg(mi, 4), xmi, ni, i0, D, S_img, mod_stride, d.eps, st));
      QIMG_TRY(qimg_ln_modulate_rows(xt, seg(mt, 3), seg(mt, 4), xmt, nt, t0, D, T, mod_stride, d.eps, st));
      {
        qimg_gemm_problem p[2];
        memset(p, 0, sizeof p);
        own_problem(p[0], xmi, bw.img_mlp_w1, bw.img_mlp_b1, ni, FF, D, S_img, i0);
        p[0].out = h_img; p[0].ldo = FF;
        own_problem(p[1], xmt, bw.txt_mlp_w1, bw.txt_mlp_b1, nt, FF, D, T, t0);
        p[1].out = h_txt; p[1].ldo = FF;
        QIMG_TRY(qimg_gemm(p, nt > 0 ? 2 : 1, QIMG_EPI_BIAS_GELU, st));
      }
      {
        qimg_gemm_problem p[2];
        memset(p, 0, sizeof p);
        own_problem(p[0], h_img, bw.img_mlp_w2, bw.img_mlp_b2, ni, D, FF, S_img, i0);
        p[0].out = xi; p[0].ldo = D; p[0].gate = seg(mi, 5); p[0].gate_stride = mod_stride;
        own_problem(p[1], h_txt, bw.txt_mlp_w2, bw.txt_mlp_b2, nt, D, FF, T, t0);
        p[1].out = xt; p[1].ldo = D; p[1].gate = seg(mt, 5); p[1].gate_stride = mod_stride;
        QIMG_TRY(qimg_gemm(p, nt > 0 ? 2 : 1, QIMG_EPI_BIAS_GATE_RES, st));
      }
    }
    // epilogue on my rows; the [rows, out_dim] result (0.5 MB per image) is all-gathered into every rank's workspace
    QIMG_TRY(qimg_ln_modulate_rows(xi, (const char*)emb_out + (size_t)D * 2, emb_out, xmi, ni, i0, D, S_img, emb_stride, d.eps, st));
    {
      qimg_gemm_problem p;
      memset(&p, 0, sizeof p);
      p.A = xmi; p.W = e->g.proj_out_w; p.bias = e->g.proj_out_b; p.M = ni; p.N = d.out_dim; p.K = D;
      p.rows_per_batch = S_img; p.row_base = i0; p.out = ws + w.out_own; p.ldo = d.out_dim;
      QIMG_TRY(qimg_gemm(&p, 1, QIMG_EPI_BIAS, st));
    }
    {
      void* dst[8];
      for (int o = 0; o < sp; ++o) dst[o] = (char*)e->peer_ws[o] + w.out_all + (size_t)i0 * d.out_dim * 2;
      QIMG_TRY(tp_p2p_push_rows(ws + w.out_own, dst, (long long)ni * d.out_dim, sp, (cudaStream_t)st));
    }
    QIMG_TRY(tp_p2p_barrier(e->peer_flags, sp, r, (cudaStream_t)st));
    QIMG_CUDA_CHECK(cudaMemcpyAsync(out, ws + w.out_all, (size_t)Mi * d.out_dim * 2, cudaMemcpyDeviceToDevice, (cudaStream_t)st));
    // the next forward's first pushes (q / k / v of block 0) must not overtake a peer still copying out_all: it reads only
    // its OWN workspace there, and nobody writes out_all before the next forward's final push, which lies behind 2 L barriers
    return 0;
  }

  // ---- 60 dual-stream blocks -----------------------------------------------------------------
  const float sm_scale = 1.0f / sqrtf(128.0f);
  struct PredicateScope {  // the BLOCKS stage runs under the step cache's device predicate (if any)
    explicit PredicateScope(const int* f) { set_launch_predicate(f); }
    ~PredicateScope() { set_launch_predicate(nullptr); }
  } predicate_scope((stages & QIMG_STAGE_BLOCKS) ? e->blocks_predicate : nullptr);
  if (e->blocks_predicate && (stages & QIMG_STAGE_BLOCKS) && tp > 1)
    return fail("qimg_engine_forward_stages: a blocks predicate is not supported under tensor parallelism");
  for (int l = 0; (stages & QIMG_STAGE_BLOCKS) && l < L; ++l) {
    QIMG_RANGE("qimg.block");
    const qimg_block_weights& bw = e->blocks[l];
    // modulation layout per stream: [shift1, scale1, gate1, shift2, scale2, gate2] x D  (chunk(2) then chunk(3))
    const char* mi = mod_all + ((size_t)l * 12 * D) * 2;
    const char* mt = mi + (size_t)6 * D * 2;
    auto seg = [&](const char* base, int i) { return (const void*)(base + (size_t)i * D * 2); };

    const bool p2p = tp > 1 && e->p2p;
    if (!p2p || l == 0) {  // peer-memory TP: blocks > 0 receive their LN1 output from the previous block's fused reduction
      QIMG_TRY(qimg_ln_modulate(x_img, seg(mi, 0), seg(mi, 1), xm_img, Mi, D, S_img, mod_stride, d.eps, st));
      QIMG_TRY(qimg_ln_modulate(x_txt, seg(mt, 0), seg(mt, 1), xm_txt, Mt, D, T, mod_stride, d.eps, st));
    }
    {
      QIMG_RANGE("gemm.qkv+qknorm+rope");
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = xm_img; p[0].W = bw.to_qkv_w; p[0].bias = bw.to_qkv_b; p[0].M = Mi; p[0].N = 3 * Dl; p[0].K = D;
      p[0].rows_per_batch = S_img; p[0].q = q; p[0].k = k; p[0].v = v; p[0].norm_q_w = bw.norm_q; p[0].norm_k_w = bw.norm_k;
      p[0].rope_cos = img_cos; p[0].rope_sin = img_sin; p[0].S_joint = S; p[0].pos_off = T; p[0].H = Hl; p[0].eps = d.eps;
      p[1] = p[0];
      p[1].A = xm_txt; p[1].W = bw.add_kv_w; p[1].bias = bw.add_kv_b; p[1].M = Mt; p[1].rows_per_batch = T;
      p[1].norm_q_w = bw.norm_added_q; p[1].norm_k_w = bw.norm_added_k; p[1].rope_cos = txt_cos; p[1].rope_sin = txt_sin;
      p[1].pos_off = 0;
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_QKV, st));
    }
    {
      QIMG_RANGE("fmha.joint");
      QIMG_TRY(qimg_fmha_joint(q, k, v, at_txt, at_img, B, Hl, S, T, sm_scale, st));
    }
    QIMG_RANGE("gemm.out_proj / ln2 / gemm.mlp_up / gemm.mlp_down (+ tp reductions)");
    if (tp == 1) {
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = at_img; p[0].W = bw.to_out_w; p[0].bias = bw.to_out_b; p[0].M = Mi; p[0].N = D; p[0].K = D;
      p[0].rows_per_batch = S_img; p[0].out = x_img; p[0].ldo = D; p[0].gate = seg(mi, 2); p[0].gate_stride = mod_stride;
      p[1] = p[0];
      p[1].A = at_txt; p[1].W = bw.to_add_out_w; p[1].bias = bw.to_add_out_b; p[1].M = Mt; p[1].rows_per_batch = T;
      p[1].out = x_txt; p[1].gate = seg(mt, 2);
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS_GATE_RES, st));
    } else if (p2p) {
      // row-parallel out-projection fused with the reduce-scatter: fp32 partial tiles go straight to the row owners
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = at_img; p[0].W = bw.to_out_w; p[0].M = Mi; p[0].N = D; p[0].K = Dl; p[0].rows_per_batch = S_img;
      partial_problem(p[0], 0);
      p[1] = p[0];
      p[1].A = at_txt; p[1].W = bw.to_add_out_w; p[1].M = Mt; p[1].rows_per_batch = T;
      partial_problem(p[1], w.own_img);
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_PARTIAL_F32, st));
      QIMG_TRY(tp_reduce_ln(bw.to_out_b, seg(mi, 2), bw.to_add_out_b, seg(mt, 2), mod_stride, seg(mi, 3), seg(mi, 4), seg(mt, 3),
                            seg(mt, 4), mod_stride));
    } else {
      // row-parallel out-projection: partial sums over the local heads -> all-reduce -> bias + gate + residual
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = at_img; p[0].W = bw.to_out_w; p[0].bias = zero_bias; p[0].M = Mi; p[0].N = D; p[0].K = Dl;
      p[0].rows_per_batch = S_img; p[0].out = part; p[0].ldo = D;
      p[1] = p[0];
      p[1].A = at_txt; p[1].W = bw.to_add_out_w; p[1].M = Mt; p[1].rows_per_batch = T; p[1].out = part_txt;
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS, st));
      QIMG_TRY(tp_reduce(bw.to_out_b, seg(mi, 2), bw.to_add_out_b, seg(mt, 2), mod_stride));
    }
    if (!p2p) {
      QIMG_TRY(qimg_ln_modulate(x_img, seg(mi, 3), seg(mi, 4), xm_img, Mi, D, S_img, mod_stride, d.eps, st));
      QIMG_TRY(qimg_ln_modulate(x_txt, seg(mt, 3), seg(mt, 4), xm_txt, Mt, D, T, mod_stride, d.eps, st));
    }
    {
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = xm_img; p[0].W = bw.img_mlp_w1; p[0].bias = bw.img_mlp_b1; p[0].M = Mi; p[0].N = FFl; p[0].K = D;
      p[0].rows_per_batch = S_img; p[0].out = h_img; p[0].ldo = FFl;
      p[1] = p[0];
      p[1].A = xm_txt; p[1].W = bw.txt_mlp_w1; p[1].bias = bw.txt_mlp_b1; p[1].M = Mt; p[1].rows_per_batch = T; p[1].out = h_txt;
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS_GELU, st));
    }
    if (tp == 1) {
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = h_img; p[0].W = bw.img_mlp_w2; p[0].bias = bw.img_mlp_b2; p[0].M = Mi; p[0].N = D; p[0].K = FF;
      p[0].rows_per_batch = S_img; p[0].out = x_img; p[0].ldo = D; p[0].gate = seg(mi, 5); p[0].gate_stride = mod_stride;
      p[1] = p[0];
      p[1].A = h_txt; p[1].W = bw.txt_mlp_w2; p[1].bias = bw.txt_mlp_b2; p[1].M = Mt; p[1].rows_per_batch = T;
      p[1].out = x_txt; p[1].gate = seg(mt, 5);
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS_GATE_RES, st));
    } else if (p2p) {
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = h_img; p[0].W = bw.img_mlp_w2; p[0].M = Mi; p[0].N = D; p[0].K = FFl; p[0].rows_per_batch = S_img;
      partial_problem(p[0], 0);
      p[1] = p[0];
      p[1].A = h_txt; p[1].W = bw.txt_mlp_w2; p[1].M = Mt; p[1].rows_per_batch = T;
      partial_problem(p[1], w.own_img);
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_PARTIAL_F32, st));
      if (l + 1 < L) {  // the LayerNorm that follows is block l+1's first one
        const char* ni = mod_all + ((size_t)(l + 1) * 12 * D) * 2;
        const char* nt = ni + (size_t)6 * D * 2;
        QIMG_TRY(tp_reduce_ln(bw.img_mlp_b2, seg(mi, 5), bw.txt_mlp_b2, seg(mt, 5), mod_stride, seg(ni, 0), seg(ni, 1), seg(nt, 0),
                              seg(nt, 1), mod_stride));
      } else {  // ... or norm_out (AdaLayerNormContinuous: scale first, then shift); the text stream ends here
        const void* sc = emb_out;
        const void* sh = (const char*)emb_out + (size_t)D * 2;
        QIMG_TRY(tp_reduce_ln(bw.img_mlp_b2, seg(mi, 5), bw.txt_mlp_b2, seg(mt, 5), mod_stride, sh, sc, sh, sc, emb_stride));
      }
    } else {
      qimg_gemm_problem p[2];
      memset(p, 0, sizeof p);
      p[0].A = h_img; p[0].W = bw.img_mlp_w2; p[0].bias = zero_bias; p[0].M = Mi; p[0].N = D; p[0].K = FFl;
      p[0].rows_per_batch = S_img; p[0].out = part; p[0].ldo = D;
      p[1] = p[0];
      p[1].A = h_txt; p[1].W = bw.txt_mlp_w2; p[1].M = Mt; p[1].rows_per_batch = T; p[1].out = part_txt;
      QIMG_TRY(qimg_gemm(p, 2, QIMG_EPI_BIAS, st));
      QIMG_TRY(tp_reduce(bw.img_mlp_b2, seg(mi, 5), bw.txt_mlp_b2, seg(mt, 5), mod_stride));
    }
  }

  // ---- epilogue: AdaLayerNormContinuous (scale first, then shift) + proj_out -------------------
  if (stages & QIMG_STAGE_POST) {
  QIMG_RANGE("qimg.post (norm_out, proj_out)");
  if (!(tp > 1 && e->p2p))  // peer-memory TP: the last block's fused reduction already produced norm_out's rows in xm_img
    QIMG_TRY(qimg_ln_modulate(x_img, (const char*)emb_out + (size_t)D * 2, emb_out, xm_img, Mi, D, S_img, emb_stride, d.eps, st));
  {
    qimg_gemm_problem p;
    memset(&p, 0, sizeof p);
    p.A = xm_img; p.W = e->g.proj_out_w; p.bias = e->g.proj_out_b; p.M = Mi; p.N = d.out_dim; p.K = D;
    p.rows_per_batch = S_img; p.out = out; p.ldo = d.out_dim;
    QIMG_TRY(qimg_gemm(&p, 1, QIMG_EPI_BIAS, st));
  }
  }  // QIMG_STAGE_POST
  return 0;
}

}  // extern "C"
