/* pixelrec_b200 -- C ABI of the B200-native SASRec hot path (libpixelrec_b200.so).
 *
 * The reference (westlake-repl/PixelRec) has NO native / FFI boundary: its hot path is stock
 * PyTorch ops inside the Python model plugins.  This ABI is therefore defined by this build and
 * sits one level beneath the Python plugin classes (pixelrec_b200/model/...), each entry point
 * replacing one group of implicit ATen/cuBLAS kernels of the reference.  The file:line after
 * "replaces" is relative to /root/reference/code.
 *
 * Conventions (all functions):
 *   - plain C types; every pointer is a DEVICE pointer unless the name ends in _host;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); work is enqueued
 *     asynchronously on it, nothing synchronises the host;
 *   - nothing allocates or frees caller memory; scratch comes from the caller, sized by the
 *     matching *_workspace_bytes() query;
 *   - return 0 on success, PR_ERR_INVALID_ARGUMENT (<0) for a rejected argument (message in
 *     pr_last_error_string(), thread-local), or a positive cudaError_t from a failed launch;
 *   - fp32 tensors are row-major and contiguous unless a stride is part of the signature;
 *     feature dims must be multiples of 4 floats (16-byte rows) -- true for every reference config;
 *   - re-entrant across streams and devices (no global mutable state besides cached
 *     per-device attributes and the process-wide switches pr_set_tuning / pr_set_seed_device).
 */
#ifndef PIXELREC_B200_H_
#define PIXELREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PR_ABI_VERSION 1
#define PR_OK 0
#define PR_ERR_INVALID_ARGUMENT (-1)
#define PR_ERR_UNSUPPORTED (-2)

#if defined(__GNUC__)
#define PR_API __attribute__((visibility("default")))
#else
#define PR_API
#endif

typedef void* pr_stream_t;

/* activation ids for pr_act_* (REC/model/layers.py:640-648 ACT2FN) */
#define PR_ACT_GELU 0    /* erf form, layers.py:651-660 */
#define PR_ACT_RELU 1
#define PR_ACT_SWISH 2
#define PR_ACT_TANH 3
#define PR_ACT_SIGMOID 4
#define PR_ACT_QUICK_GELU 5   /* x * sigmoid(1.702 x): CLIP ViT item encoder (REC/model/load.py:91-99, HF CLIPVisionModel) */

PR_API int pr_version(void);
PR_API const char* pr_last_error_string(void);
/* number of SMs of the current device (148 on B200); <0 on error */
PR_API int pr_sm_count(void);
/* Makes `device` current for this library's CUDA runtime on the calling thread.  The host side calls
 * it with the device of the tensors it passes (one process per GPU: once, with LOCAL_RANK). */
PR_API int pr_set_device(int device);
/* Kernel-variant switches for A/B measurement (bit mask; also read once from the environment variable PR_TUNE):
 *   1 = LayerNorm backward as per-warp bulk-copy row pipelines, 2 = L2 prefetch of the next row in the register LN kernels,
 *   4 = LayerNorm forward as per-warp bulk-copy row pipelines, 8 = tensor-core attention with two warps per item pipeline,
 *   16 = score_topk with the branch-free 8-warp epilogue, 32 = (with 16) table tile TMA-multicast across a cluster,
 *   128 = fp16 scoring keeps the 128 x D seq_out tile resident in shared memory (D <= 512),
 *   64 = long-sequence attention (forward; backward for dh <= 64) on tensor cores (TF32 operands: results differ from the
 *        fp32 kernels within TF32 tolerance),
 *   256 = table-gradient segment reduce as a TMA-staged shared-memory ring (rows of 256 B .. 8 KiB) instead of the LDG
 *         warp-per-run kernel, 512 = LayerNorm forward (register kernel) with two rows in flight per warp.
 * mask < 0 only queries.  Returns the mask in effect.  Results are identical under every mask except where noted. */
PR_API int pr_set_tuning(int mask);
/* CUDA-graph replay of a training step (staged): every dropout kernel adds *seed_offset_dev (a uint64 in device memory, bumped
 * by a kernel captured in the same graph) to the host-side seed frozen into the graph, so that replay j draws the masks the
 * eager step j would have drawn.  NULL switches back to host seeds.  Only builds compiled with -DPR_SEED_DEV implement it
 * (PR_BUILD_DEFS=-DPR_SEED_DEV python -m pixelrec_b200.build --force); the default build returns PR_ERR_UNSUPPORTED for a
 * non-NULL pointer and its kernels are byte-identical to a build without this entry point. */
PR_API int pr_set_seed_device(const uint64_t* seed_offset_dev);

/* ------------------------------------------------------------------------------------------
 * K1  embedding row gather.     replaces nn.Embedding.forward: REC/model/IDNet/sasrec.py:31,68
 *                               (same call in gru4rec.py:25,51)
 *   out[r, :] = W[idx[r], :]   r in [0,R).  Bit-exact copy, pad row 0 included.
 *   status (optional, may be NULL): *status |= 1 if any idx is outside [0,N) (that row is
 *   zero-filled) -- the reference raises IndexError / device-asserts there.
 *   impl: 0 = auto, 1 = LDG.128/STG.128 path, 2 = TMA bulk-copy (smem-staged) path.
 */
PR_API int pr_gather_rows_f32(const float* W, int64_t N, int64_t D, const int64_t* idx, int64_t R, float* out,
                       int32_t* status, int impl, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2  gradient scatter-add.     replaces autograd's embedding_dense_backward of sasrec.py:68
 *                               + optimizer.zero_grad's dense fill (trainer/trainer.py:117,122)
 * Two calls so the index work (depends only on idx) can be done once and overlap the forward:
 *
 * pr_scatter_plan: stable LSD radix sort of (idx, position), segment boundaries, compaction.
 *   perm[R]        positions r sorted by (idx[r], r)
 *   uniq_ids[U]    ascending distinct ids, padding_idx removed     (capacity min(R,N))
 *   seg_start[U+1] perm offsets of each id's run                   (capacity min(R,N)+1)
 *   n_uniq[1]      U
 *   row2slot       optional [N]: row2slot[uniq_ids[u]] = u; caller pre-fills with -1 once,
 *                  pr_adamw_rows_f32 restores the -1s it consumes
 *   status         optional: |= 1 on out-of-range idx (such rows are dropped)
 *
 * pr_scatter_add_rows_f32: out_rows[u, :] = scale * sum_{k in run u} dOut[perm[k], :], added
 *   sequentially in ascending position (deterministic; oracle/sasrec_np.py scatter_add_rows
 *   defines the same order, so the comparison is bit-exact for scale == 1).
 *   If dense_G != NULL the same rows are also stored to dense_G[uniq_ids[u], :] (no zero fill).
 *   R <= 2^24 rows per call.
 */
PR_API size_t pr_scatter_plan_workspace_bytes(int64_t R, int64_t N);
PR_API int pr_scatter_plan(const int64_t* idx, int64_t R, int64_t N, int64_t padding_idx, int32_t* perm, int32_t* uniq_ids,
                    int32_t* seg_start, int32_t* n_uniq, int32_t* row2slot, void* workspace, size_t workspace_bytes,
                    int32_t* status, pr_stream_t stream);
PR_API int pr_scatter_add_rows_f32(const float* dOut, int64_t R, int64_t D, const int32_t* perm, const int32_t* uniq_ids,
                            const int32_t* seg_start, const int32_t* n_uniq, int64_t max_uniq, float scale,
                            float* out_rows, float* dense_G, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K10 AdamW.                    replaces torch.optim.AdamW.step: trainer/trainer.py:100-103,125
 * Exact dense semantics (every row decays every step) with a SPARSE gradient:
 *   g(row) = grad_scale * grad_rows[row2slot[row]] if row2slot[row] >= 0 else 0.
 * step >= 1 is the 1-based step count used for bias correction; if step_dev != NULL the count is
 * read from device memory instead (CUDA-graph friendly).  row2slot entries are reset to -1.
 * pr_adamw_dense_f32: the same update over a flat dense buffer (all non-table parameters).
 */
PR_API int pr_adamw_rows_f32(float* W, float* M, float* V, int64_t N, int64_t D, const float* grad_rows, int32_t* row2slot,
                      float lr, float beta1, float beta2, float eps, float weight_decay, float grad_scale,
                      int64_t step, const int64_t* step_dev, pr_stream_t stream);
PR_API int pr_adamw_dense_f32(float* w, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                       float eps, float weight_decay, float grad_scale, int64_t step, const int64_t* step_dev,
                       pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K3/K7  (dropout) + residual add + LayerNorm (+ dropout).
 *   replaces sasrec.py:77-83 (pos-emb add, LayerNorm, dropout)            [post-LN dropout]
 *        and layers.py:613-615, 669-671 (out_dropout, LayerNorm(h + x))   [pre-add dropout]
 *   z[r,:] = drop_pre(h[r,:]) + res[res_row(r),:] ;  y = drop_post(LN(z) * gamma + beta)
 *   h row r = (s,t), s = r / rows_per_seq, t = r % rows_per_seq lives at h + s*h_seq_stride + t*D
 *   (lets the kernel read input_emb = item_emb[:,0,:-1] straight out of the [B,2,L+1,D] gather
 *   output without a copy).  res_period > 0: res_row(r) = r % res_period (position embedding
 *   broadcast over the batch), else res_row(r) = r.
 *   Dropout masks: Philox4x32-10(seed; counter, stream id), 16-bit fields, keep iff field >= round(p*2^16) (layout in ln.cu),
 *   kept values scaled by 1/(1-p); p == 0 disables.  mean/rstd [rows] are saved for backward.
 * Backward returns dh (grad of h, dropout applied; row (s,t) at dh + s*dh_seq_stride + t*D, or
 * contiguous [rows,D] when dh_seq_stride == 0; dh_accumulate != 0 adds into dh instead of storing --
 * lets the embedding LayerNorm add straight into the [B,2,L+1,D] table-gradient rows), dres (grad
 * of z, un-reduced: [rows,D]; NULL to skip), and per-CTA partial column sums for dgamma/dbeta: partials [2, n_partials, D] which
 * pr_colsum_f32 reduces deterministically.
 */
PR_API int pr_add_ln_fwd_f32(const float* h, int64_t h_seq_stride, int64_t rows_per_seq, const float* res, int64_t res_period,
                      const float* gamma, const float* beta, float eps, int64_t rows, int64_t D, float p_pre,
                      float p_post, uint64_t seed, uint32_t stream_pre, uint32_t stream_post, float* y, float* mean,
                      float* rstd, pr_stream_t stream);
PR_API int pr_add_ln_bwd_partials(int64_t rows, int64_t D);
PR_API int pr_add_ln_bwd_f32(const float* dy, const float* h, int64_t h_seq_stride, int64_t rows_per_seq, const float* res,
                      int64_t res_period, const float* gamma, const float* mean, const float* rstd, int64_t rows,
                      int64_t D, float p_pre, float p_post, uint64_t seed, uint32_t stream_pre, uint32_t stream_post,
                      float* dh, int64_t dh_seq_stride, int dh_accumulate, float* dres, float* partials,
                      int n_partials, pr_stream_t stream);
/* Same as pr_add_ln_bwd_f32 plus a third partial matrix: column sums of dh = bias gradient of the Linear that produced h
 * (layers.py:613 `dense`, :669 `dense_2`): partials [3, n_partials, D] = (dgamma, dbeta, dbias).  Saves one full re-read
 * of dh per Linear (ATen's sum(0) in the autograd path). */
PR_API int pr_add_ln_bwd_bias_f32(const float* dy, const float* h, int64_t h_seq_stride, int64_t rows_per_seq, const float* res,
                      int64_t res_period, const float* gamma, const float* mean, const float* rstd, int64_t rows,
                      int64_t D, float p_pre, float p_post, uint64_t seed, uint32_t stream_pre, uint32_t stream_post,
                      float* dh, int64_t dh_seq_stride, int dh_accumulate, float* dres, float* partials,
                      int n_partials, pr_stream_t stream);
/* Backward of LayerNorm(z) where z = drop_pre(h) + res was WRITTEN by the producing Linear (pr_gemm_tf32_drop: bias + dropout +
 * residual in the GEMM epilogue, layers.py:613-615 / 669-671), so the forward is pr_add_ln_fwd_f32(z, res = NULL, p_pre = 0) and
 * the backward reads two tensors instead of three:  dz = LN'(dy) (gradient of the residual branch; NULL when p_pre == 0, where
 * dh == dz),  dh = drop_pre mask applied to dz (same Philox draws as the epilogue),  partials [3, n_partials, D] = (dgamma,
 * dbeta, column sums of dh = bias gradient of the Linear).  Contiguous [rows, D] tensors. */
PR_API int pr_add_ln_bwd_bias_z_f32(const float* dy, const float* z, const float* gamma, const float* mean, const float* rstd,
                      int64_t rows, int64_t D, float p_pre, uint64_t seed, uint32_t stream_pre, float* dh, float* dz,
                      float* partials, int n_partials, pr_stream_t stream);
/* out[m, c] = sum_p partials[m, p, c]   (m < n_mats, p < n_partials), fixed order */
PR_API int pr_colsum_f32(const float* partials, int n_mats, int n_partials, int64_t D, float* out, pr_stream_t stream);

/* out[c] = sum_r x[r, c] for a dense row-major [M, C] matrix (C % 4 == 0): the bias gradient of the fused q|k|v projection
 * (layers.py:586-588), whose output gradient comes from the attention backward.  partials: [pr_colsum_rows_partials(M, C), C]. */
PR_API int pr_colsum_rows_partials(int64_t M, int64_t C);
PR_API int pr_colsum_rows_f32(const float* x, int64_t M, int64_t C, float* partials, int n_partials, float* out, pr_stream_t stream);

/* activation of the feed-forward layer: layers.py:640-660,667   y = act(x) ; dx = act'(x) * dy */
PR_API int pr_act_fwd_f32(const float* x, int64_t n, int act, float* y, pr_stream_t stream);
PR_API int pr_act_bwd_f32(const float* x, const float* dy, int64_t n, int act, float* dx, pr_stream_t stream);
/* row-wise variant that also emits per-CTA partial column sums of dx [n_partials, cols] (bias grad of dense_1, layers.py:666) */
PR_API int pr_act_bwd_bias_partials(int64_t rows, int64_t cols);
PR_API int pr_act_bwd_bias_f32(const float* x, const float* dy, int64_t rows, int64_t cols, int act, float* dx, float* partials,
                        int n_partials, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4+K6  causal self-attention core.   replaces get_attention_mask sasrec.py:119-126 and
 *                                      MultiHeadAttention.forward layers.py:590-612
 *   q,k,v: row (b,t) at ptr + (b*L + t)*ld, head hd occupies floats [hd*dh, (hd+1)*dh)
 *   key_ids [B,L] int64: key j is valid iff key_ids[b,j] != 0 (masked_index in training,
 *   item_seq in predict); NULL = all valid.  causal != 0 adds j <= i.
 *   S = q k^T / sqrt(dh) + (valid ? 0 : -1e9); P = softmax(S); ctx = drop(P) v
 *   ctx [B,L,h*dh] (heads re-concatenated, layers.py:610-612); probs [B,h,L,L] saved (pre-dropout).
 *   Fully-masked query rows give the reference's uniform 1/L row (never NaN).
 *   Limits: L <= 64, dh % 4 == 0, dh <= 128 or dh % 128 == 0.
 */
PR_API int pr_sasrec_attn_fwd_f32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids, int B,
                           int L, int h, int dh, int causal, float p_drop, uint64_t seed, uint32_t rng_stream,
                           float* ctx, float* probs, pr_stream_t stream);
PR_API int pr_sasrec_attn_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const float* probs,
                           const float* dctx, int B, int L, int h, int dh, int causal, float p_drop, uint64_t seed,
                           uint32_t rng_stream, float* dq, float* dk, float* dv, int64_t ld_grad, pr_stream_t stream);

/* Tensor-core variant of the attention core (same arguments, same saved tensors, same dropout masks): the two
 * per-head products run as mma.sync TF32 with fp32 accumulation, tiles arrive by 2-D swizzled TMA.  Limits: L <= 32, dh % 32 == 0.  Used when TF32
 * matrix products are allowed (as for the linear layers); the _f32 entry points above are the strict-fp32 path. */
PR_API int pr_sasrec_attn_fwd_tf32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids, int B,
                            int L, int h, int dh, int causal, float p_drop, uint64_t seed, uint32_t rng_stream,
                            float* ctx, float* probs, pr_stream_t stream);
PR_API int pr_sasrec_attn_bwd_tf32(const float* q, const float* k, const float* v, int64_t ld, const float* probs,
                            const float* dctx, int B, int L, int h, int dh, int causal, float p_drop, uint64_t seed,
                            uint32_t rng_stream, float* dq, float* dk, float* dv, int64_t ld_grad, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K8  sampled-negative pairwise loss.   replaces sasrec.py:88-92 (== gru4rec.py:63-67, mosasrec.py:89-93)
 *   pos[b,t] = <out[b,t], tp[b,t]>, neg[b,t] = <out[b,t], tn[b,t]>,
 *   loss = mean_b( -sum_t log(sigmoid(pos-neg) + 1e-8) * mask[b,t] )
 *   tp/tn row (b,t) at ptr + b*t_seq_stride + t*D (reads item_emb[:,0,1:] / [:,1,1:] in place).
 *   Saves coef[b,t] = dloss/d(pos-neg) (incl. the 1/B of the mean) for backward; loss_terms [B,L]
 *   scratch (per-position terms, summed in a fixed order -> deterministic); loss [1].
 *   pos_score / neg_score [B,L] optional (NULL to skip); positions with mask == 0 are not read.
 * Backward: d_out = g*coef*(tp - tn), d_tp = g*coef*out, d_tn = -g*coef*out, g = *dloss (device scalar).
 *   d_tp/d_tn rows are written with d_seq_stride (so they can land inside a [B,2,L+1,D] buffer).
 */
PR_API int pr_bpr_loss_fwd_f32(const float* out, const float* tp, const float* tn, int64_t t_seq_stride,
                        const int64_t* mask, int64_t B, int64_t L, int64_t D, float* pos_score, float* neg_score,
                        float* coef, float* loss_terms, float* loss, pr_stream_t stream);
PR_API int pr_bpr_loss_bwd_f32(const float* out, const float* tp, const float* tn, int64_t t_seq_stride, const float* coef,
                        const float* dloss, int64_t B, int64_t L, int64_t D, float* d_out, float* d_tp, float* d_tn,
                        int64_t d_seq_stride, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K9  full-catalog scoring + mask + top-k, tcgen05 tensor cores (TF32 in, fp32 accumulate in TMEM).
 *   replaces  scores = seq_output @ item_feature.T           REC/model/IDNet/sasrec.py:112 (gru4rec.py:79)
 *             scores[:,0] = -inf; scores[hist_u,hist_i]=-inf REC/trainer/trainer.py:334-336
 *             torch.topk(scores, max(topk))                  REC/evaluator/collector.py:133
 *   seq_out [B_e, D], W [N, D] (item_feature), D % 32 == 0; hist_u/hist_i [n_hist] int64 (may be NULL when
 *   n_hist == 0); mask_col0 != 0 masks the [PAD] column.  Outputs topk_val [B_e, k] (descending; -inf when
 *   fewer than k unmasked items) and topk_idx [B_e, k] int64 (ties: lower item id first; -1 for -inf slots).
 *   The [B_e, N] score matrix is never written.  k <= 32.  workspace: 256-byte aligned device buffer.
 */
PR_API size_t pr_score_topk_workspace_bytes(int64_t B_e, int64_t N, int k);
PR_API int pr_score_topk_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D,
                             const int64_t* hist_u, const int64_t* hist_i, int64_t n_hist, int mask_col0, int k,
                             float* topk_val, int64_t* topk_idx, void* workspace, size_t workspace_bytes,
                             pr_stream_t stream);

/* K9, id-exact: the ranking the reference computes on fp32 scores (collector.py:133).  Same fused TF32 pass as above as a
 *   candidate generator (the 32 best of the row's per-thread lists), then every candidate is re-scored in fp32 and ranked (ties:
 *   lower item id); a row whose k-th fp32 score is not provably above everything outside its candidates -- bound: the largest
 *   TF32 value any list dropped or the 32nd candidate, + 1.25 * 2^-9 * |seq_out[row]| * max_j |W[j]| -- is re-ranked over the
 *   whole catalog in fp32 (rare; *n_fallback_rows, a device int32, counts them; may be NULL).  k <= 16, N >= 32.  w_norm_max: device float holding max_j |W[j]| (pr_table_norm_max_f32;
 *   compute it once per evaluation), or NULL to have it computed here.  topk_val are fp32 scores.
 */
PR_API size_t pr_score_topk_exact_workspace_bytes(int64_t B_e, int64_t N, int k);
PR_API int pr_table_norm_max_f32(const float* W, int64_t N, int64_t D, float* out_max, pr_stream_t stream);
PR_API int pr_score_topk_exact_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D,
                                   const int64_t* hist_u, const int64_t* hist_i, int64_t n_hist, int mask_col0, int k,
                                   const float* w_norm_max, float* topk_val, int64_t* topk_idx, int32_t* n_fallback_rows,
                                   void* workspace, size_t workspace_bytes, pr_stream_t stream);

/* K5: the encoder's linear layers AND their backward on tcgen05 (csrc/gemm.cu): one persistent CTA-pair kernel
 *   (tcgen05.mma.cta_group::2, UMMA 256x256x8 TF32, TMEM double-buffered accumulators, TMA-fed 5-stage ring, TMA-store epilogue).
 *   replaces nn.Linear forward, REC/model/layers.py:586-588, :613, :666, :669, and autograd's two GEMMs per Linear.
 *     out[m, n] = epilogue( sum_k A(m, k) * B(n, k) ),   m < M, n < N, k < K;  fp32 in memory, TF32 operands, fp32 accumulate.
 *   A: a_mn == 0 -> K-major, memory [M][K] with row stride lda;  a_mn != 0 -> MN-major, memory [K][M] with row stride lda.
 *   B: b_mn == 0 -> K-major, memory [N][K] (nn.Linear weight layout), row stride ldb;  b_mn != 0 -> memory [K][N].
 *       forward      y  = x W^T + b :  A = x  (K-major),  B = W  (K-major)
 *       input grad   dx = dy W      :  A = dy (K-major),  B = W  (MN-major: memory [out_features][in_features] = [K][N])
 *       weight grad  dW = dy^T x    :  A = dy (MN-major: memory [rows][out] = [K][M]),  B = x (MN-major), splits > 1
 *   epi: PR_GEMM_STORE    out = acc + bias
 *        PR_GEMM_ADD      out = acc + bias + aux                      (aux [M, N]: residual / gradient accumulation)
 *        PR_GEMM_ACT      out = act(acc + bias); out2 (optional) = acc + bias    (dense_1 + gelu, layers.py:666-667)
 *        PR_GEMM_ACT_BWD  out = acc * act'(aux)                       (aux = the pre-activation: input grad through the activation)
 *   bias [N] may be NULL.  colsum_partials (optional, [pr_gemm_colsum_rows(M), N], no bias): per-32-row column sums of `out`
 *   (bias gradient; reduce with pr_colsum_f32).  splits > 1 (PR_GEMM_STORE without bias only): `out` is [splits, M, N], slab s holding
 *   the partial sum over k-blocks [s*ceil(K/32/splits), ...); pr_gemm_splitk_reduce_f32 adds the slabs in ascending order.
 *   flags bit 0 (PR_GEMM_DEBIAS): the tensor core reads fp32 operands as TF32 by TRUNCATING the low 13 mantissa bits: an operand
 *   with mantissa m in [1, 2) loses on average 2^-11 / m of its value, i.e. 0.72 * 2^-11 = 3.5e-4 for log-uniform mantissas, so
 *   every product is ~7.1e-4 too small (measured on Gaussian operands: -7.16e-4) -- a systematic shrink per GEMM that compounds
 *   through a chain of layers (measured: -0.29 % on a weight gradient) where round-to-nearest errors would average out.  With
 *   the flag the accumulator is multiplied by 1.00071 before the epilogue: the error becomes zero-mean with the variance of
 *   round-to-nearest TF32.  Leave it off for operands that are exactly representable in TF32 (nothing is truncated then).
 *   N % 4 == 0, K % 4 == 0 for a K-major operand (a K tail is zero-filled by TMA), M % 4 == 0 for an MN-major A, all pointers
 *   16-byte aligned, out / out2 / aux dense [M, N].
 */
#define PR_GEMM_DEBIAS 1
#define PR_GEMM_STORE 0
#define PR_GEMM_ADD 1
#define PR_GEMM_ACT 2
#define PR_GEMM_ACT_BWD 3
PR_API int pr_gemm_colsum_rows(int64_t M);
PR_API int pr_gemm_tf32(const float* A, int a_mn, int64_t lda, const float* B, int b_mn, int64_t ldb, int64_t M, int64_t N,
                        int64_t K, const float* bias, const float* aux, int epi, int act, float* out, float* out2, int splits,
                        float* colsum_partials, int flags, pr_stream_t stream);
/* PR_GEMM_ADD with the hidden dropout of layers.py:614 / :670 in the epilogue:  out = drop(acc + bias) + aux, the keep bits being
 *   exactly those pr_add_ln_fwd_f32 would draw for (seed, rng_stream) on a [M, N] tensor (Philox4x32-10, layout in ln.cu), so that
 *   pr_add_ln_bwd_bias_z_f32 and the oracle (oracle/philox_np.py) regenerate the same mask.  One draw serves 8 outputs.  p_drop in [0, 1). */
PR_API int pr_gemm_tf32_drop(const float* A, int a_mn, int64_t lda, const float* B, int b_mn, int64_t ldb, int64_t M, int64_t N,
                             int64_t K, const float* bias, const float* aux, float* out, int flags, float p_drop, uint64_t seed,
                             uint32_t rng_stream, pr_stream_t stream);
PR_API int pr_gemm_splitk_reduce_f32(const float* partials, int splits, int64_t n, float* out, pr_stream_t stream);

/* K9 with fp16 operands (staged): fp16 has the 10 explicit mantissa bits of TF32 (and is rounded to nearest, where the TF32
 *   datapath reads truncated fp32 words) but kind::f16 MMAs run at twice the TF32 rate on half the operand bytes.  The
 *   exponent range is narrower: |x| > 65504 saturates and raises status bit 2; |x| < 6e-5 loses precision (absolute error
 *   <= 3e-8).  LayerNorm outputs and N(0, 0.02^2)-scale tables are far inside that range.
 * pr_score_prepare_f16: dst_f16[i] = fp16(src[i]), n % 4 == 0 -- convert the item table ONCE per evaluation.
 * pr_score_topk_f16: as pr_score_topk_f32 with W16 = the converted [N, D] table (D % 64 == 0); seq_out is fp32 and converted
 *   into the workspace.  Always runs the v2 kernel (pr_set_tuning bit 32 adds the cluster multicast).
 */
PR_API int pr_score_prepare_f16(const float* src, int64_t n, void* dst_f16, int32_t* status, pr_stream_t stream);
PR_API size_t pr_score_topk_f16_workspace_bytes(int64_t B_e, int64_t N, int64_t D, int k);
PR_API int pr_score_topk_f16(const float* seq_out, int64_t B_e, const void* W16, int64_t N, int64_t D, const int64_t* hist_u,
                             const int64_t* hist_i, int64_t n_hist, int mask_col0, int k, float* topk_val, int64_t* topk_idx,
                             void* workspace, size_t workspace_bytes, int32_t* status, pr_stream_t stream);

/* full-catalog softmax cross-entropy on the same tcgen05 pipeline (north_star: "scoring ... fused with the softmax/CE").
 *   EXTENSION -- the reference trains with sampled negatives (REC/model/IDNet/sasrec.py:88-92) and has no full-softmax
 *   loss; the oracle is a restatement of F.cross_entropy(seq_out @ W.T, target) (oracle/sasrec_np.py full_catalog_ce).
 *   For every row: lse = log sum_c exp(<seq_out[row], W[c]>) over the catalog (column 0 excluded when mask_col0 != 0),
 *   tgt_logit = <seq_out[row], W[target[row]]>, nll = lse - tgt_logit.  Any of lse / tgt_logit / nll may be NULL (not all);
 *   target may be NULL when only lse is wanted.  The logits are never written (online max / sum in the GEMM epilogue).
 *   Forward; pr_ce_grad_chunk_f32 below is the backward's building block.
 */
PR_API size_t pr_score_ce_workspace_bytes(int64_t B_e, int64_t N);
PR_API int pr_score_ce_f32(const float* seq_out, int64_t B_e, const float* W, int64_t N, int64_t D, const int64_t* target,
                           int mask_col0, float* lse, float* tgt_logit, float* nll, void* workspace, size_t workspace_bytes,
                           pr_stream_t stream);

/* Backward building block of the same extension (pixelrec_b200/ops.py ScoreCEFn walks the catalog in column chunks: recompute a
 *   chunk of logits with pr_gemm_tf32, turn it into dS with this kernel, then dX += dS W_c and dW_c = dS^T X on pr_gemm_tf32, so the
 *   [B_e, N] probabilities are never held either).  In place on S [rows, C] (row stride ld), columns = items c0 .. c0 + C - 1:
 *   S[r, j] <- dnll[r] * (exp(S[r, j] - lse[r]) - [c0 + j == target[r]]);  the padding column (item 0) <- 0 when mask_col0 != 0. */
PR_API int pr_ce_grad_chunk_f32(float* S, int64_t ld, int64_t rows, int64_t C, int64_t c0, const float* lse, const int64_t* target,
                                const float* dnll, int mask_col0, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * A1  on-device batch construction.   replaces SEQTrainDataset.__getitem__ + default collate,
 *                                     REC/data/dataset/trainset.py:40-75 (python loops in 10 DataLoader workers)
 *   padded [n_seq, W] int64: training windows left-padded with 0 (W = MAX_ITEM_LIST_LENGTH + 1), resident in HBM;
 *   sel [B] int64: which windows form this batch.  Outputs items [B,2,W] (positives | aligned uniform negatives
 *   rejected against the sequence's own items, 0 where there is no transition) and mask [B,W-1] (masked_index).
 *   Deterministic in (seed, b, t): Philox4x32-10, see csrc/sampler.cu.  status bit 0: a sel index out of range.
 */
PR_API int pr_seq_batch_build(const int64_t* padded, int64_t n_seq, int W, const int64_t* sel, int64_t B, int64_t item_num,
                       uint64_t seed, int64_t* items, int64_t* mask, int32_t* status, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6b attention core for LONG sequences, 64 < L <= 256 (any L <= 256 is accepted), strict fp32.
 *   replaces  the attention core of the ViT item encoder when it has more than 64 tokens (ViT-B/16: 197 tokens, 12 heads,
 *             dh 64): HF CLIPVisionModel encoder layers built by REC/model/load.py:90-99, used at
 *             REC/model/PixelNet/mosasrec.py:69; and REC/model/layers.py:590-612 for MAX_ITEM_LIST_LENGTH > 64.
 *   Same operand layout and mask semantics as pr_sasrec_attn_*_f32 (fused q|k|v rows with leading dimension ld, key_ids
 *   optional, additive -1e9 mask); no dropout (the CLIP encoder has none).  dh must be a power of two in [4, 128].
 *   Outputs ctx [B, L, h*dh] and lse [B*h, L] (log-sum-exp of the masked, scaled scores): the [L, L] probabilities are
 *   never written; the backward recomputes them from lse.  delta_ws: [B*h, L] fp32 scratch.
 */
PR_API int pr_attn_long_fwd_f32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids, int B, int L,
                                int h, int dh, int causal, float* ctx, float* lse, pr_stream_t stream);
PR_API int pr_attn_long_bwd_f32(const float* q, const float* k, const float* v, int64_t ld, const int64_t* key_ids,
                                const float* ctx, const float* lse, const float* dctx, int B, int L, int h, int dh, int causal,
                                float* dq, float* dk, float* dv, int64_t ld_grad, float* delta_ws, pr_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Peer-memory exchange of the ROW-SHARDED item table (one process per GPU, NVLink / NVSwitch P2P).
 *   replaces  the table replicated per GPU + the dense [N,D] gradient all-reduce of DDP, REC/run.py:40 around
 *             nn.Embedding REC/model/IDNet/sasrec.py:31,68; SURVEY.md section 8e (owner(i) = i % G, local row i / G).
 * pr_shared_alloc / pr_shared_free: device memory of the CURRENT device that other processes may map (cudaMalloc +
 *   CUDA IPC); handle64 receives an opaque 64-byte handle to ship to the peers (e.g. all_gather_object).
 * pr_shared_open / pr_shared_close: map / unmap a peer's allocation into the current device's address space (peer access
 *   is enabled on demand).  Call it in a process OTHER than the one that allocated; open each handle once per process.
 * pr_gather_rows_peers_f32: out[r, :] = shards[idx[r] % G][idx[r] / G, :] -- lookup and exchange in one kernel; shards is
 *   a DEVICE array of G device pointers (own shard included), N the global row count.  status bit 0: idx outside [0,N).
 * pr_push_rows_peers_f32: for every u < U with ids[u] != skip_id: claims slot p = counters[ids[u] % G]++ and writes
 *   rows[u, :] to recv_rows[owner][rank*cap + p, :] and ids[u] / G to recv_ids[owner][rank*cap + p].  recv_rows[g] is
 *   rank g's receive buffer [G*cap, D], recv_ids[g] its id buffer [G*cap] (the owner pre-fills it with -1 = unused);
 *   counters [G] int32 must be zero on entry.  ids must be distinct.  status bit 1: a region overflowed (row dropped).
 *   The caller orders these kernels against the owners' use of the buffers (a barrier collective on the stream).
 */
PR_API int pr_shared_alloc(size_t bytes, void** dptr, unsigned char* handle64);
PR_API int pr_shared_free(void* dptr);
PR_API int pr_shared_open(const unsigned char* handle64, void** dptr);
PR_API int pr_shared_close(void* dptr);
PR_API int pr_gather_rows_peers_f32(const float* const* shards, int G, int64_t N, int64_t D, const int64_t* idx, int64_t R,
                                    float* out, int32_t* status, pr_stream_t stream);
PR_API int pr_push_rows_peers_f32(const float* rows, const int64_t* ids, int64_t U, int64_t D, int G, int rank, int64_t cap,
                                  int64_t skip_id, float* const* recv_rows, int64_t* const* recv_ids, int32_t* counters,
                                  int32_t* status, pr_stream_t stream);
/* Barrier over peer-mapped flags: flag_tables[r] = device address (mapped here) of rank r's uint64 flag array [G], all zero at
 * start; `epoch` = 1, 2, 3, ... the same sequence on every rank.  Returns (on the stream) once every rank has called it with this
 * epoch; everything rank r wrote into peer memory before its call is visible to kernels launched after it.  status bit 4 = a
 * peer did not arrive within ~2 s. */
PR_API int pr_peer_barrier(uint64_t* const* flag_tables, int G, int rank, uint64_t epoch, uint64_t* epoch_dev, int32_t* status,
                           pr_stream_t stream);
/* epoch_dev != NULL: the call's epoch is ++(*epoch_dev) taken on the device (`epoch` ignored), so a captured CUDA graph can be
 * replayed.
 *
 * Device-side exchange plan: the peer kernels driven by ONE pr_scatter_plan of the step's ids (padding id dropped), the number
 * of distinct ids staying in device memory -- no host synchronisation, the multi-GPU step is CUDA-graph capturable.
 *   pr_plan_inverse                 inverse[r] = slot u of request position r in uniq_ids, pad_slot for dropped positions
 *   pr_gather_rows_peers_plan_f32   out[u] = owner's row of uniq_ids[u], u < *n_uniq; out[pad_slot] = row of pad_id (pad_slot >= 0);
 *                                   out has max_uniq + 1 rows; expand to request order with pr_gather_rows_f32(out, inverse)
 *   pr_push_rows_peers_plan_f32     pr_push_rows_peers_f32 over ids[0 .. *n_dev) (int32 global ids) */
PR_API int pr_plan_inverse(const int32_t* perm, const int32_t* seg_start, const int32_t* n_uniq, int64_t R, int64_t pad_slot,
                           int64_t* inverse, pr_stream_t stream);
PR_API int pr_gather_rows_peers_plan_f32(const float* const* shards, int G, int64_t N, int64_t D, const int32_t* uniq_ids,
                                         const int32_t* n_uniq, int64_t max_uniq, int64_t pad_id, int64_t pad_slot, float* out,
                                         int32_t* status, pr_stream_t stream);
PR_API int pr_push_rows_peers_plan_f32(const float* rows, const int32_t* ids, const int32_t* n_dev, int64_t max_n, int64_t D, int G,
                                       int rank, int64_t cap, float* const* recv_rows, int64_t* const* recv_ids, int32_t* counters,
                                       int32_t* status, pr_stream_t stream);


#ifdef __cplusplus
}
#endif
#endif /* PIXELREC_B200_H_ */
