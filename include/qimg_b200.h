/* qimg_b200.h — C ABI of the B200 (sm_100a) Qwen-Image DiT denoising engine.
 *
 * This is the drop-in boundary: plain pointers + sizes + a cudaStream_t,
 * no torch types.  Every device pointer is bf16 unless stated; the CALLER owns all
 * allocations (torch tensors on the reference side); the library owns only a process-wide
 * TMA-descriptor cache.  Every function returns 0 on success, non-zero on error, with a
 * message available from qimg_last_error() — the Python layer raises RuntimeError so the
 * reference's worker error path (diffusion/worker/gpu_worker.py:267-274) keeps working.
 * There is no CPU fallback: on a machine without an sm_100 device calls fail loudly.
 *
 * Each entry point names the reference interface (vllm-omni @ be81443, file:line) it replaces.
 */
#ifndef QIMG_B200_H
#define QIMG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* qimg_stream_t; /* cudaStream_t */

/* ---- library / device --------------------------------------------------------------- */
int qimg_abi_version(void);
const char* qimg_last_error(void);
/* 0 iff the current CUDA device is compute capability 10.x (B200/GB200); fills sm count. */
int qimg_device_check(int* sm_count);
/* number of kernels this library launched since the last reset (bench.py "gpu_launches") */
long long qimg_launch_count(void);
void qimg_reset_launch_count(void);

/* Per-launch CUDA-event timing of the tensor-core kernels (kind 0 = tcgen05 GEMM, 1 = FMHA): while
 * enabled every launch is bracketed by events on its stream; collect() synchronises them and returns the
 * summed device time, launch count and algorithmic FLOPs since the previous collect (bench.py roofline). */
void qimg_prof_enable(int on);
int qimg_prof_collect(int kind, double* ms_total, long long* launches, double* flops_total);

/* NVTX ranges around the engine's launch families ("qimg.forward" > "qimg.pre" / "qimg.block" > "gemm.qkv+qknorm+rope",
 * "fmha.joint", ... / "qimg.post"); default off, env QIMG_NVTX=1 turns them on.  Header-only NVTX3: free without a profiler. */
void qimg_set_nvtx(int on);

/* GEMM tile mode: 0 = one CTA per 128x256 tile (tcgen05 cta_group::1), 1 = CTA pair per 256x256 tile
 * (cta_group::2, cluster of 2), 2 (default) = the pair unless 128-row tiles save whole rounds of the grid (small M: one
 * image split over several GPUs).  Results are bit-identical across modes.  Env QIMG_GEMM_MODE overrides. */
int qimg_set_gemm_mode(int mode);
int qimg_get_gemm_mode(void);
/* Raster band height in 256-row tiles (default 8; env QIMG_GEMM_GROUP_M): tiles of a band visit every weight column while
 * the band's activations stay in L2; each band streams the weight matrix from HBM once. */
int qimg_set_gemm_group_m(int tiles);

/* Attention pipeline, mode = pipeline | (poly << 3):
 *   pipeline 4 = EXACT: every KV tile's row maximum is reduced (and exchanged between the two threads of a row) before
 *                the tile is exponentiated; correct for any input;
 *            6 = FAST (default): tile j >= 1 is exponentiated in one pass against the running reference maximum the row
 *                already has, the tile's own maximum only decides the lazy rebase of O / l afterwards, P is handed to the
 *                tensor pipe in two halves.  Exact as long as no score exceeds the reference by more than 2^100 (a jump
 *                of > 69 nats inside one 128-key tile); such a launch sets a device-side flag — see qimg_fmha_overflow —
 *                and the caller must recompute with pipeline 4 (the native denoise loop does, once per 50 steps).
 *   poly 1 = 25 % of the softmax exponentials on a degree-3 FMA-pipe polynomial instead of MUFU.EX2.
 * Env QIMG_FMHA_MODE overrides the default.  (Round 1 shipped seven pipelines; five lost and were deleted.) */
int qimg_set_fmha_mode(int mode);
int qimg_get_fmha_mode(void);
/* Query tiling of the attention grid: 0 = one CTA per PAIR of 128-row query tiles (K/V fetched once per 256 rows),
 * 1 = one CTA per tile, -1 (default) = per tile while tiles x B x H still fits ONE wave of SMs (e.g. the 3 local heads of an
 * 8-way split at B = 1: 99 CTAs instead of 51), pairs otherwise.  Each row's result does not depend on the tiling. */
int qimg_set_fmha_single_tile(int mode);

/* ---- bandwidth-bound fused ops ------------------------------------------------------ */
/* y[r,:] = LN(x[r,:]; eps, no affine) * (1 + scale[b,:]) + shift[b,:],  b = r / rows_per_batch.
 * Replaces AdaLayerNorm.forward_cuda/forward_native, vllm_omni/diffusion/layers/adalayernorm.py:62-68,94-102
 * (and AdaLayerNormContinuous at qwen_image_transformer.py:797).  mod_stride = elements between
 * consecutive batches of shift/scale (0 => one modulation row shared by the whole batch). */
int qimg_ln_modulate(const void* x, const void* shift, const void* scale, void* y, int rows, int D, int rows_per_batch,
                     long long mod_stride, float eps, qimg_stream_t stream);
/* Per-token modulation select — AdaLayerNorm.preprocess / forward with `index` (layers/adalayernorm.py:31-54,94-102; the
 * reference builds the index for `zero_cond_t` models, qwen_image_transformer.py:748-754): shift / scale hold 2 * index_batch
 * modulation rows, token r (int32 index[r], device) of batch b uses row b when index[r] == 0 and row index_batch + b otherwise.
 * qimg_select_rows gathers the matching per-token gate rows: out[r, :] = src[(index[r] ? index_batch : 0) + r / rows_per_batch, :]. */
int qimg_ln_modulate_indexed(const void* x, const void* shift, const void* scale, void* y, int rows, int D, int rows_per_batch,
                             long long mod_stride, float eps, const int* index, int index_batch, qimg_stream_t stream);
int qimg_select_rows(const void* src, long long src_stride, const int* index, void* out, int rows, int D, int rows_per_batch,
                     int index_batch, qimg_stream_t stream);
/* Same on a slice of the stream: x / y point at the slice, row_base is the global index of its first row (selects the
 * batch's modulation row).  Sequence-parallel engine mode. */
int qimg_ln_modulate_rows(const void* x, const void* shift, const void* scale, void* y, int rows, int row_base, int D,
                          int rows_per_batch, long long mod_stride, float eps, qimg_stream_t stream);

/* x[r,:] += gate[b,:] * y[r,:]  — gated residual, qwen_image_transformer.py:586-587,592,597. */
int qimg_gate_residual(void* x, const void* y, const void* gate, int rows, int D, int rows_per_batch,
                       long long gate_stride, qimg_stream_t stream);

/* Tensor-parallel form of the gated residual: y is the ALL-REDUCED partial sum of a row-parallel linear
 * (to_out / net.2 sharded over K), bias is added after the reduction:  x += gate * (y + bias). */
int qimg_gate_residual_bias(void* x, const void* y, const void* bias, const void* gate, int rows, int D, int rows_per_batch,
                            long long gate_stride, qimg_stream_t stream);

/* y = RMSNorm(x; w, eps) over the last dim (vLLM RMSNorm used as txt_norm, qwen_image_transformer.py:669,758). */
int qimg_rms_norm(const void* x, const void* w, void* y, int rows, int D, float eps, qimg_stream_t stream);

/* y[M, N] (leading dim ldy) = act(x[M,K]) @ W[N,K]^T + bias, small M (<= 64); act_silu != 0 applies
 * SiLU to x first.  Replaces nn.Sequential(SiLU, Linear) img_mod/txt_mod (qwen_image_transformer.py:478-481,
 * 494-497,552-557 — all 2*L projections in one call when W is the concatenated weight), the
 * TimestepEmbedding MLP (:45,50-62) and norm_out.linear (:686). */
int qimg_linear_small_m(const void* x, const void* W, const void* bias, void* y, int M, long long N, int K,
                        long long ldy, int act_silu, qimg_stream_t stream);

/* out[B,256] = [cos | sin](1000 * t * f) — diffusers Timesteps(256, flip_sin_to_cos=True, shift 0, scale 1000),
 * qwen_image_transformer.py:44 (restated in-tree at pipeline_qwen_image.py:135-184).  t is bf16 [B]. */
int qimg_timestep_sinusoid(const void* t, void* out, int B, qimg_stream_t stream);

/* Fused true-CFG combine + norm rescale + flow-match Euler step on latents [rows, 64]:
 *   pipeline_qwen_image.py:580-583 and FlowMatchEulerDiscreteScheduler.step at :585.
 * neg == NULL => no CFG.  latents updated in place: x = bf16(float(x) + bf16(bf16(sigma_next - sigma) * noise))
 * (torch promotes the 0-dim fp32 dt to the bf16 operand dtype). */
int qimg_cfg_euler_step(const void* pos, const void* neg, void* latents, long long rows, int C, float cfg_scale,
                        float sigma, float sigma_next, qimg_stream_t stream);
/* Same step with (sigma_i, sigma_{i+1}) read from DEVICE memory (fp32 [2]): the launch has no per-timestep host argument,
 * so one captured CUDA graph of a denoise step can be replayed for all timesteps (pipeline `enable_cuda_graph`). */
int qimg_cfg_euler_step_dev(const void* pos, const void* neg, void* latents, long long rows, int C, float cfg_scale,
                            const float* sigma_pair, qimg_stream_t stream);
/* How `dt * model_output` treats dt: 0 (default) = rounded to bf16 first — torch's result for a 0-dim fp32 DEVICE tensor
 * times a bf16 tensor, which is what diffusers' FlowMatchEulerDiscreteScheduler.step computes with its sigmas on the
 * device, and what torch computes on CPU; 1 = kept at fp32 (torch's result when dt is a 0-dim CPU tensor and the
 * model output is on CUDA). */
int qimg_set_euler_dt_fp32(int on);

/* ---- tcgen05 GEMM family ------------------------------------------------------------ */
enum { QIMG_EPI_BIAS = 0, QIMG_EPI_BIAS_GELU = 1, QIMG_EPI_BIAS_GATE_RES = 2, QIMG_EPI_QKV = 3, QIMG_EPI_PARTIAL_F32 = 4 };

/* One linear problem out = A[M,K] @ W[N,K]^T (+ epilogue).  Up to two problems (image stream,
 * text stream) are grouped in one launch.  Replaces vLLM QKVParallelLinear/ReplicatedLinear and
 * diffusers FeedForward linears at qwen_image_transformer.py:380-385,452-456,491,501 together with
 * the pointwise ops the epilogue absorbs (see csrc/qimg_gemm.cuh). */
typedef struct qimg_gemm_problem {
  const void* A;    /* [M, K] row-major, K % 8 == 0 */
  const void* W;    /* [N, K] row-major (nn.Linear weight) */
  const void* bias; /* [N] */
  int M, N, K;
  int rows_per_batch; /* rows per image of this stream (gate / position lookup) */
  /* QIMG_EPI_BIAS / _GELU: out [M, ldo].  QIMG_EPI_BIAS_GATE_RES: out is the residual stream x, updated in
   * place: x = x + gate[b,:] * (A W^T + bias). */
  void* out;
  int ldo;
  const void* gate;
  long long gate_stride;
  /* QIMG_EPI_QKV (N = 3*H*128): per-head RMSNorm(q,k) + interleaved RoPE, scattered into joint
   * head-major q/k/v [B, H, S_joint, 128] at sequence offset pos_off. */
  void* q;
  void* k;
  void* v;
  const void* norm_q_w;
  const void* norm_k_w;
  const void* rope_cos; /* [rows_per_batch, 64] bf16 */
  const void* rope_sin;
  int S_joint, pos_off, H;
  float eps;
  /* QIMG_EPI_PARTIAL_F32 (tensor-parallel row-parallel linear, K sharded over tp_size ranks; bias is NOT applied):
   * the fp32 partial sums of row m are stored into the receive buffer of the rank that owns m — rows are split
   * contiguously and balanced (the first M % tp_size owners get one more) — at
   *   (float*)tp_recv[owner] + ((tp_rank * tp_recv_rows + tp_recv_row_off + (m - first_row(owner))) * N + n).
   * tp_recv[] are device pointers valid in THIS process (peer mappings opened with qimg_ipc_open_handle, the own
   * buffer at index tp_rank).  The reduction itself is qimg_tp_reduce_ln_push. */
  void* tp_recv[8];
  int tp_size, tp_rank, tp_recv_rows, tp_recv_row_off;
  /* Sequence parallelism: A holds a rank's OWN rows of the stream; row_base is the global index of its first row (batch,
   * position and gate lookups are done on global rows; `out` / A are addressed with local rows).  With sp_size > 1 the
   * QIMG_EPI_QKV epilogue stores head h into the q/k/v buffers of the rank that owns the head — sp_q/k/v[h / (H / sp_size)],
   * local head h % (H / sp_size), each [B, H / sp_size, S_joint, 128] — i.e. the all-to-all in front of Ulysses attention
   * (reference attention/parallel/ulysses.py:110-112) as peer stores of the GEMM epilogue; q / k / v are then unused. */
  int row_base;
  int sp_size;
  void* sp_q[8];
  void* sp_k[8];
  void* sp_v[8];
} qimg_gemm_problem;

int qimg_gemm(const qimg_gemm_problem* problems, int nprob, int epilogue, qimg_stream_t stream);

/* ---- joint attention ---------------------------------------------------------------- */
/* softmax(Q K^T * scale) V over the joint [text; image] sequence, non-causal, no mask.
 * q,k,v: [B, H, S, 128] head-major (as written by QIMG_EPI_QKV); out_txt [B*T, H*128], out_img
 * [B*(S-T), H*128].  Replaces Attention.forward -> SDPAImpl.forward (attention/layer.py:54-70,
 * backends/sdpa.py:46-66) and the cat/split around it (qwen_image_transformer.py:414-416,448-449). */
int qimg_fmha_joint(const void* q, const void* k, const void* v, void* out_txt, void* out_img, int B, int H, int S,
                    int T, float softmax_scale, qimg_stream_t stream);
/* Same with an explicit pipeline (`mode` as in qimg_set_fmha_mode; < 0 = the process default).  The layer-level
 * AttentionImpl plug-in passes 4 (exact): it has no end-of-denoise point at which to consult the overflow flag. */
int qimg_fmha_joint_mode(const void* q, const void* k, const void* v, void* out_txt, void* out_img, int B, int H, int S,
                         int T, float softmax_scale, int mode, qimg_stream_t stream);
/* Sequence-parallel (Ulysses) form: this rank holds H_local heads of q / k / v over ALL S rows (written there by every
 * rank's QKV epilogue, qimg_gemm_problem.sp_q/k/v) and the output rows are stored straight into the buffers of the ranks
 * that OWN them: out_img[o] / out_txt[o] = owner o's [own rows, H_local * sp_size * 128] buffers (device pointers valid in
 * this process); rows of each stream are split contiguously and balanced over the sp_size owners.  The two all-to-alls of
 * reference attention/parallel/ulysses.py:27-135 become peer stores of the producing kernels. */
typedef struct qimg_fmha_sp {
  int sp_size, sp_rank;
  void* out_img[8];
  void* out_txt[8];
} qimg_fmha_sp;
int qimg_fmha_joint_sp(const void* q, const void* k, const void* v, int B, int H_local, int S, int T, float softmax_scale,
                       const qimg_fmha_sp* sp, qimg_stream_t stream);
/* Overflow flag of the fast pipeline on the current device: *out = 1 if any launch since the last reset saw a score more
 * than 2^100 above its row's reference maximum (results of that launch are then not trustworthy).  Synchronising 4-byte
 * read; reset != 0 clears the flag.  Reference semantics being guarded: exact softmax, attention/backends/sdpa.py:56-64. */
int qimg_fmha_overflow(int* out, int reset);
/* Diagnostics: when set to a device buffer of 32 int64, the attention kernels record the cycle
 * counters of CTA 200: [0] MMA-warp loop, [1..4] its waits on K, P1, V, P0, [5] KV tiles; [8..14] / [16..22] tile-0 /
 * tile-1 softmax warp: loop, wait on S, score load (+ max in pipeline 4), pair barrier, rebase check,
 * exponentials + P stores, tail (store wait, fence, arrive).  NULL disables.  Needs a -DQIMG_FMHA_TRACE build. */
int qimg_set_fmha_trace(void* dev_buf_32_i64);

/* ---- whole-model engine ------------------------------------------------------------- */
typedef struct qimg_dims {
  int num_layers, num_heads, head_dim /* must be 128 */, in_channels /* 64 */, out_dim /* 64 */, joint_dim /* 3584 */;
  float eps;
} qimg_dims;

typedef struct qimg_block_weights { /* names: SURVEY.md §8c / qwen_image_transformer.py:804-839 */
  const void *img_mod_w, *img_mod_b, *txt_mod_w, *txt_mod_b;     /* [6D, D], [6D] */
  const void *to_qkv_w, *to_qkv_b, *add_kv_w, *add_kv_b;         /* [3D, D], [3D] */
  const void *norm_q, *norm_k, *norm_added_q, *norm_added_k;     /* [128] */
  const void *to_out_w, *to_out_b, *to_add_out_w, *to_add_out_b; /* [D, D], [D] */
  const void *img_mlp_w1, *img_mlp_b1, *img_mlp_w2, *img_mlp_b2; /* [4D, D], [4D], [D, 4D], [D] */
  const void *txt_mlp_w1, *txt_mlp_b1, *txt_mlp_w2, *txt_mlp_b2;
} qimg_block_weights;

typedef struct qimg_global_weights {
  const void *t_lin1_w, *t_lin1_b, *t_lin2_w, *t_lin2_b; /* time_text_embed.timestep_embedder.linear_{1,2} */
  const void* txt_norm_w;
  const void *img_in_w, *img_in_b, *txt_in_w, *txt_in_b;
  const void *norm_out_w, *norm_out_b, *proj_out_w, *proj_out_b;
  /* Optional: all 2*L modulation weights concatenated as [L][img,txt][6D, D] / [L][2][6D]; when
   * non-NULL the engine computes every block's modulation with ONE small-M launch per forward. */
  const void *mod_all_w, *mod_all_b;
} qimg_global_weights;

typedef struct qimg_engine qimg_engine;

/* Tensor parallelism over attention heads / FFN (not present in the reference, whose Qwen-Image
 * linears are all `disable_tp=True`, qwen_image_transformer.py:318-349).  With tp_size P the caller passes
 * per-rank weight shards: to_qkv/add_kv_proj rows of the local H/P heads ([3D/P, D], q|k|v), to_out/to_add_out
 * columns of the local heads ([D, D/P]), MLP up rows / down columns of the local FF/P slice; biases of the
 * row-parallel linears (to_out, to_add_out, net.2) are passed in full and applied once after the reduction.
 * Everything else is replicated.  The engine calls `allreduce(buf, count, user, stream)` (sum over the TP
 * group, in place, `count` bf16 elements, enqueued on `stream`) twice per block: after the attention
 * out-projection and after the MLP down-projection (image and text partial sums share one buffer). */
typedef int (*qimg_allreduce_fn)(void* buf, long long count, void* user, qimg_stream_t stream);
int qimg_engine_set_tp(qimg_engine* e, int tp_size, qimg_allreduce_fn allreduce, void* user);

/* Peer-memory tensor parallelism (one node, NVLink/NVSwitch): instead of the all-reduce callback + epilogue
 * kernel, each rank reduces ITS slice of rows straight out of every rank's partial-sum buffer (P2P loads, fp32
 * accumulation, one rounding), applies bias + gate + residual and stores the new residual rows into every
 * rank's workspace (P2P stores) — reduce-scatter + epilogue + all-gather in one kernel, bracketed by two
 * cross-GPU flag barriers.  No NCCL on the data path.
 *   qimg_p2p_alloc            cudaMalloc'ed, zero-filled buffer (IPC-exportable, unlike a sub-allocation of a
 *                             caching allocator); the workspace and a >= 128-byte flag buffer come from here
 *   qimg_ipc_get_handle       64-byte cudaIpcMemHandle_t of such a buffer, to be sent to the peer processes
 *   qimg_ipc_open_handle      maps a peer's buffer into this process (enables peer access lazily)
 *   qimg_engine_set_tp_p2p    peer_workspaces[p] / peer_flags[p] for p in [0, tp_size): this rank's own
 *                             (local) pointers at index tp_rank, the opened peer mappings elsewhere.
 *                             qimg_engine_forward must then be given peer_workspaces[tp_rank] as workspace,
 *                             and all ranks must issue the same sequence of forwards.
 *   qimg_engine_p2p_error     synchronising read of the barrier time-out flag (0 = healthy). */
int qimg_p2p_alloc(size_t bytes, void** out);
int qimg_p2p_free(void* ptr);
int qimg_ipc_get_handle(const void* dev_ptr, void* handle64);
int qimg_ipc_open_handle(const void* handle64, void** out);
int qimg_ipc_close_handle(void* ptr);
int qimg_engine_set_tp_p2p(qimg_engine* e, int tp_size, int tp_rank, void* const* peer_workspaces, void* const* peer_flags);
/* Sequence parallelism (the reference's own multi-GPU mode for this model: Ulysses, attention/parallel/ulysses.py:27-135,
 * `ulysses_degree`), fused: every rank keeps the FULL weights (like a data-parallel replica), runs the LayerNorms and all
 * four linears of a block on ITS rows only, and attention on ITS heads over all rows.  The two all-to-alls are not separate
 * collectives: the QKV GEMM epilogue stores each head's q / k / v rows into the head owner's buffers and the attention
 * epilogue stores each output row into the row owner's buffer (peer stores over NVLink), with one cross-GPU flag barrier
 * after each.  Per block and rank (P - 1) / P * own rows * (3 D + D) * 2 bytes leave the GPU — about a tenth of the
 * tensor-parallel mode — and every dot product runs over its full K on one GPU, so the result is BIT-IDENTICAL to the
 * single-GPU engine.  Same registration protocol as qimg_engine_set_tp_p2p (NULL arrays = declare the mode). */
int qimg_engine_set_sp_p2p(qimg_engine* e, int sp_size, int sp_rank, void* const* peer_workspaces, void* const* peer_flags);
int qimg_engine_p2p_error(qimg_engine* e, int* out);

int qimg_engine_create(const qimg_dims* dims, const qimg_global_weights* g, const qimg_block_weights* blocks,
                       qimg_engine** out);
void qimg_engine_destroy(qimg_engine* e);
size_t qimg_engine_workspace_bytes(const qimg_engine* e, int B, int S_img, int T);

/* One DiT forward = QwenImageTransformer2DModel.forward (qwen_image_transformer.py:692-802), SP off:
 *   hidden [B,S_img,64], enc [B,T,joint], timestep bf16 [n_t] (already /1000; n_t = B, or 1 when the
 *   whole batch shares one timestep as in QwenImagePipeline.diffuse, pipeline_qwen_image.py:552),
 *   RoPE tables bf16 img [S_img,64] x2, txt [T,64] x2  ->  out [B,S_img,64].
 * workspace: >= qimg_engine_workspace_bytes(), 1024-byte aligned, caller-owned. */
int qimg_engine_forward(qimg_engine* e, const void* hidden, const void* enc, const void* timestep, int n_t,
                        const void* img_cos, const void* img_sin, const void* txt_cos, const void* txt_sin, int B,
                        int S_img, int T, void* out, void* workspace, size_t workspace_bytes, qimg_stream_t stream);

/* Staged forward for step caches (TeaCache, reference cache/teacache/hook.py:80-165 + extractors.py:184-246):
 *   QIMG_STAGE_PRE     img_in, txt_norm + txt_in, temb, all modulations (extractor preprocessing :187-204); when the
 *                      blocks are NOT part of the same call, block 0's modulated image stream (:206-209) is also left
 *                      in the workspace at qimg_engine_ws_offset_mod()
 *   QIMG_STAGE_BLOCKS  the L dual-stream blocks on the residual streams held in the workspace (run_transformer_blocks)
 *   QIMG_STAGE_POST    norm_out + proj_out -> out (postprocess :230-236)
 * The residual streams live at qimg_engine_ws_offset_img/_txt between calls; the workspace must not be reused by another
 * shape in between.  qimg_engine_forward == all three stages. */
#define QIMG_STAGE_PRE 1
#define QIMG_STAGE_BLOCKS 2
#define QIMG_STAGE_POST 4
#define QIMG_STAGE_ALL 7
int qimg_engine_forward_stages(qimg_engine* e, int stages, const void* hidden, const void* enc, const void* timestep, int n_t,
                               const void* img_cos, const void* img_sin, const void* txt_cos, const void* txt_sin, int B,
                               int S_img, int T, void* out, void* workspace, size_t workspace_bytes, qimg_stream_t stream);
size_t qimg_engine_ws_offset_mod(const qimg_engine* e, int B, int S_img, int T);

/* Step-cache arithmetic on bf16 tensors of n elements (n % 8 == 0), each a single pass:
 *   qimg_rel_l1_sums      sums2[0] = sum |bf16(a - b)|, sums2[1] = sum |b| (fp32, device memory; zeroed by the call):
 *                         the two means of hook.py:198-203
 *   qimg_bf16_sub         out = bf16(a - b)       (cached residual, hook.py:152-154)
 *   qimg_bf16_add_inplace x   = bf16(x + r)       (residual reuse, hook.py:131-133) */
int qimg_rel_l1_sums(const void* a, const void* b, long long n, float* sums2, qimg_stream_t stream);
int qimg_bf16_sub(void* out, const void* a, const void* b, long long n, qimg_stream_t stream);
int qimg_bf16_add_inplace(void* x, const void* r, long long n, qimg_stream_t stream);

/* Step-cache decision ON THE DEVICE (no host synchronisation; the reference reads the distance back with .cpu().item(),
 * cache/teacache/hook.py:204-205):
 *   qimg_tea_decide     one thread: rel = bf16 arithmetic of hook.py:198-203 on sums2 / n, accum += |poly(rel)| (coef5: highest
 *                       power first, fp64 Horner = numpy.poly1d), *flag = 1 (reuse) while accum < thresh, else 0 and accum = 0.
 *                       force: 0 = decide from the data; 1 = first forward of a branch (accum = 0, compute); 2 = no previous
 *                       input yet (compute, accum kept).  hist (optional, fp32 [2 * steps]) records (flag, rel) at hist_idx.
 *   qimg_tea_residual   flag = 1: x += resid   /   flag = 0: resid = x - ori        (hook.py:131-133 / 152-154)
 *   qimg_engine_set_blocks_predicate   every kernel of the engine's BLOCKS stage starts with `if (*skip_flag) return;`, so a
 *                       reused step costs ~540 empty launches instead of 60 blocks; NULL removes the predicate. */
int qimg_tea_decide(const float* sums2, long long n, const double* coef5, double thresh, double* accum, int* flag, float* hist,
                    int hist_idx, int force, qimg_stream_t stream);
int qimg_tea_residual(void* x, const void* ori, void* resid, long long n, const int* flag, qimg_stream_t stream);
int qimg_engine_set_blocks_predicate(qimg_engine* e, const int* skip_flag);

/* Debug/test access: copies of the image / text residual streams after the last forward live in the
 * workspace at these byte offsets ([B*S_img, D] and [B*T, D] bf16). */
size_t qimg_engine_ws_offset_img(const qimg_engine* e, int B, int S_img, int T);
size_t qimg_engine_ws_offset_txt(const qimg_engine* e, int B, int S_img, int T);

/* ---- VAE decode: the post-step of QwenImagePipeline.forward ------------------------------------------
 * Replaces AutoencoderKLQwenImage._decode (vllm_omni/diffusion/models/qwen_image/autoencoder_kl_qwenimage.py:839-862; the
 * pipeline calls it at pipeline_qwen_image.py:746) for single-frame latents.  ALL pointers in this section are fp32 device
 * memory, activations are NHWC ([image, y, x, channel], `ld*` = floats between consecutive pixels).  The layer graph
 * (residual blocks, mid-block attention, up blocks) is driven by the host mirror vllm_omni_b200/.../vae_decoder.py.
 *
 * qimg_conv2d_nhwc_tf32   stride-1 "same" convolution as an implicit GEMM on tcgen05 (kind::tf32, fp32 accumulate):
 *     out[n,y,x,co] = bias[co] + res[n,y,x,co] + sum_{tap,ci} x[n, y+dy(tap), x+dx(tap), ci] * w[co, tap * Cin_pad + ci]
 *   taps = 9: 3x3 (tap = 3 * (dy + 1) + (dx + 1)); this is QwenImageCausalConv3d(k=3) on the first frame — two zero frames
 *   are padded in FRONT (:78-82), so only weight[:, :, 2] meets data — and nn.Conv2d(3, padding=1) of the resamplers (:150);
 *   Cin % 32 == 0, Cin_pad = Cin.  taps = 1: 1x1 (conv_shortcut :235, to_qkv / proj :299-300) and plain GEMMs
 *   (out[M = H*W pixels, Cout] = x[M, Cin] * w[Cout, Cin]^T: the attention's Q*K^T and P*V), Cin % 4 == 0, Cin >= 32.
 *   bias / res may be NULL.  Needs H >= 8 and W >= 16 (one 16 x 8 pixel patch per accumulator tile); Cout % 4 == 0.
 * qimg_vae_rms_act        QwenImageRMS_norm (:102-109) over the channel of each pixel, times gamma, optional SiLU (:246-247)
 * qimg_vae_upsample2x     nearest-exact x2 (QwenImageUpsample, :112-124,149)
 * qimg_vae_post_quant     post_quant_conv 1x1x1 (:848) on NCHW z [N, 16, H, W] -> NHWC [N, H, W, 32] (channels 16..31 zero:
 *                         one K block of conv_in); w [16, 16], b [16]
 * qimg_vae_conv_out       conv_out 3x3, C = 96 -> 3 (:656) on the normalised + SiLU input, clamp(-1, 1) (:857); w packed [3][9][C],
 *                         FMA pipe (exact fp32).  out (optional): NCHW fp32 [N, 3, H, W], what vae.decode returns.  out_u8
 *                         (optional): NHWC uint8 [N, H, W, 3] = the reference's post-process (VaeImageProcessor.postprocess
 *                         behind get_qwen_image_post_process_func, pipeline_qwen_image.py:40-60: (x / 2 + 0.5).clamp(0, 1)
 *                         * 255, round half to even) fused in — a quarter of the bytes for the device -> host copy
 * qimg_vae_softmax_rows   in place: s[r, :cols] = softmax(scale * s[r, :cols])   (SDPA of the mid-block attention, :321)
 * qimg_vae_transpose      out[c * rows + r] = in[r * ld_in + c] */
int qimg_conv2d_nhwc_tf32(const float* x, int ldx, const float* w, int ldw, const float* bias, const float* res, int ldr,
                          float* out, int ldo, int N, int H, int W, int Cin, int Cout, int taps, qimg_stream_t stream);
/* VAE ENCODE side (the edit pipelines' condition image, pipeline_qwen_image_edit.py:458-480 -> AutoencoderKLQwenImage._encode,
 * autoencoder_kl_qwenimage.py:793-812); it reuses the kernels above plus
 *   qimg_conv2d_down2_nhwc_tf32   the resamplers of the encoder, nn.ZeroPad2d((0, 1, 0, 1)) + nn.Conv2d(C, C, 3, stride=2)
 *                                 (:157-161): out[n, y, x, co] = bias + sum_{ky, kx, ci} x[n, 2y + ky, 2x + kx, ci] * w[co, (3 ky + kx) Cin + ci],
 *                                 input rows / columns beyond the image read as zero.  The TMA descriptor walks x and y with
 *                                 element stride 2, so the gather costs nothing.  H_in, W_in even; output [N, H_in/2, W_in/2, Cout].
 *   qimg_vae_image_to_nhwc        NCHW image [N, C <= 32, H, W] -> NHWC [N, H, W, 32], channels >= C zero (K block of conv_in) */
int qimg_conv2d_down2_nhwc_tf32(const float* x, int ldx, const float* w, int ldw, const float* bias, float* out, int ldo, int N,
                                int H_in, int W_in, int Cin, int Cout, qimg_stream_t stream);
int qimg_vae_image_to_nhwc(const float* img, float* out, int N, int C, int H, int W, qimg_stream_t stream);
/* Kernel variant of qimg_conv2d_nhwc_tf32: 0 (default) = one TMA box per (horizontal tap, channel block) shared by the three
 * vertical taps as shifted descriptor views + 16 x 16 pixel patches per CTA (2.3x fewer operand bytes per MAC);
 * 1 = one box per tap, 16 x 8 patches (the first version, kept for A/B).  Env QIMG_VAE_CONV.  Same results up to fp32
 * summation order. */
int qimg_set_vae_conv_variant(int variant);
int qimg_vae_rms_act(const float* x, const float* gamma, float* y, long long rows, int C, int silu, qimg_stream_t stream);
int qimg_vae_upsample2x(const float* x, float* out, int N, int H, int W, int C, qimg_stream_t stream);
int qimg_vae_post_quant(const float* z, const float* w, const float* b, float* out, int N, int H, int W, int z_dim,
                        qimg_stream_t stream);
int qimg_vae_conv_out(const float* x, const float* w, const float* b, float* out, uint8_t* out_u8, int N, int H, int W, int C,
                      qimg_stream_t stream);
int qimg_vae_softmax_rows(float* s, int rows, int cols, long long ld, float scale, qimg_stream_t stream);
int qimg_vae_transpose(const float* in, long long ld_in, float* out, int rows, int cols, qimg_stream_t stream);

/* ---- self test ------------------------------------------------------------------------ */
/* Raw tcgen05 probe used by tests: D[128,N] = A[128,K] B[N,K]^T through the same descriptors the
 * kernels use.  mode 0: A,B K-major from smem.  mode 1: B MN-major ([K,N] row-major in global).
 * mode 2: A staged through TMEM (tcgen05.st) as the FMHA P operand, B MN-major. */
int qimg_umma_probe(const void* A, const void* B, float* D, int N, int K, int mode, qimg_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QIMG_B200_H */
