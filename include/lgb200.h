/* lgb200 -- C ABI of the B200-native LightGlue matcher hot path.
 *
 * The reference (cvg/glue-factory) is pure Python/PyTorch: its "FFI" for this
 * path is the set of PyTorch operator calls inside
 * gluefactory/models/matchers/lightglue.py and gluefactory/models/utils/losses.py.
 * Each entry point below replaces one such group of calls; the citation on each
 * names the reference lines it stands in for.  The host-side plugin
 * (glue-factory_b200/matchers/lightglue.py) binds these with ctypes, see
 * INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers to DEVICE memory + extents + the CUDA stream to launch on;
 *    no torch types.  The caller owns every buffer; kernels keep no state.
 *  - every function returns 0 on success or a negative LGB200_ERR_* code;
 *    lgb200_last_error() returns the message of the last failure on this thread.
 *  - activations are token-major: [batch, tokens, heads, 64] for q/k/v/ctx,
 *    [batch, rows, cols] row-major for similarity/score matrices.
 *  - `dtype` is LGB200_F32 (full-precision parity path, CUDA cores) or
 *    LGB200_BF16 (bf16 operands, fp32 accumulation, tcgen05 tensor cores).
 *  - head dimension is fixed to 64 (descriptor_dim / num_heads in every
 *    LightGlue configuration, lightglue.py:349).
 */
#ifndef LGB200_H_
#define LGB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

#define LGB200_ABI_VERSION 2

#define LGB200_F32 0
#define LGB200_BF16 1

#define LGB200_OK 0
#define LGB200_ERR_INVALID (-1)     /* bad argument (null pointer, empty or misaligned input) */
#define LGB200_ERR_CUDA (-2)        /* a CUDA runtime / driver call failed */
#define LGB200_ERR_UNSUPPORTED (-3) /* shape or device not supported by this build */

int lgb200_abi_version(void);
const char* lgb200_last_error(void);
/* 0 when the current device is sm_100 (B200); LGB200_ERR_UNSUPPORTED otherwise. */
int lgb200_check_device(void);

/* ---- rotary encoding + QKV de-interleave ------------------------------------------------------
 * replaces lightglue.py:157-160 (unflatten(H,-1,3), q/k/v slices, apply_cached_rotary_emb) and
 * lightglue.py:42-49.  qkv [ntok, H*192] with feature = h*192 + d*3 + {q,k,v};
 * theta [ntok, 32] fp32 = posenc.Wr(kpts) (lightglue.py:62); q,k,v [ntok, H*64].           */
int lgb200_rope_split_fwd(const void* qkv, const float* theta, void* q, void* k, void* v, int64_t ntok, int H,
                          int dtype, cudaStream_t stream);
/* dq,dk are gradients w.r.t. the rotated q,k (saved by forward as q,k); writes dqkv [ntok, H*192]
 * and ACCUMULATES dtheta [ntok, 32] (the gradient that reaches posenc.Wr from every layer).   */
int lgb200_rope_split_bwd(const void* dq, const void* dk, const void* dv, const void* q, const void* k,
                          const float* theta, void* dqkv, float* dtheta, int64_t ntok, int H, int dtype,
                          cudaStream_t stream);

/* ---- attention ----------------------------------------------------------------------------------
 * replaces F.scaled_dot_product_attention (lightglue.py:118-121) and the cross-attention einsum/
 * softmax block (lightglue.py:207-216).  out = softmax(scale * q k^T) v per (batch, head).
 * q,out [B,Nq,H,64]; k,v [B,Nk,H,64]; lse [B,H,Nq] fp32.  Keys/values of query batch b are taken
 * from batch (b + kv_shift) % B (kv_shift = B/2 on the [image0; image1] batch = both cross
 * directions in one launch; 0 = self-attention).                                                */
int lgb200_attn_fwd(const void* q, const void* k, const void* v, void* out, float* lse, int B, int Nq, int Nk, int H,
                    int kv_shift, float scale, int dtype, cudaStream_t stream);
/* dq [B,Nq,H,64]; dk,dv [B,Nk,H,64] (indexed by KEY batch); delta_ws: lgb200_attn_bwd_ws_floats(B,Nq,Nk,H) floats of
 * scratch, 256-byte aligned (rowsum(dout*out), the per-query side data of the kernels and, for the fused bf16
 * backward, the fp32 dQ accumulator the key-tile CTAs reduce into).                                         */
int64_t lgb200_attn_bwd_ws_floats(int B, int Nq, int Nk, int H);
int lgb200_attn_bwd(const void* q, const void* k, const void* v, const void* out, const float* lse, const void* dout,
                    void* dq, void* dk, void* dv, float* delta_ws, int B, int Nq, int Nk, int H, int kv_shift,
                    float scale, int dtype, cudaStream_t stream);

/* ---- FFN LayerNorm + GELU -------------------------------------------------------------------------
 * replaces nn.LayerNorm(2D) + nn.GELU() of the ffn Sequential (lightglue.py:143-148, 178-183).
 * x,y [ntok, W], W in {256,512,1024}; mean,rstd [ntok] saved for backward.                       */
int lgb200_ln_gelu_fwd(const void* x, const float* gamma, const float* beta, void* y, float* mean, float* rstd,
                       int64_t ntok, int W, float eps, int dtype, cudaStream_t stream);
/* number of partial rows written to dgamma_part / dbeta_part / dxsum_part [parts, W] (caller sums over
 * parts); dxsum = column sums of dx = the bias gradient of the Linear that feeds the LayerNorm.  */
int lgb200_ln_gelu_bwd_parts(int64_t ntok);
int lgb200_ln_gelu_bwd(const void* dy, const void* x, const float* gamma, const float* beta, const float* mean,
                       const float* rstd, void* dx, float* dgamma_part, float* dbeta_part, float* dxsum_part,
                       int64_t ntok, int W, int dtype, cudaStream_t stream);

/* ---- batched bf16 tensor-core GEMM (tcgen05, fp32 accumulate) ---------------------------------------
 * replaces torch.einsum("bmd,bnd->bmn") of MatchAssignment (lightglue.py:283) and its two backward
 * contractions.  C[b] (M x N) = opA(A[b]) * opB(B[b]):
 *   a_mn_major = 0: A is [M,K] row-major (lda);  1: A is stored [K,M] row-major (lda)
 *   b_mn_major = 0: B is [N,K] row-major (ldb);  1: B is stored [K,N] row-major (ldb)
 * lda/ldb in elements, multiples of 8; base pointers 16-byte aligned; C row-major with ldc;
 * the result is scaled by alpha in the epilogue (the d^-1/2 of lightglue.py:282).                  */
int lgb200_gemm_bf16(const void* A, const void* B, void* C, int batch, int M, int N, int K, int a_mn_major,
                     int b_mn_major, int64_t lda, int64_t ldb, int64_t ldc, int64_t strideA, int64_t strideB,
                     int64_t strideC, int c_dtype, float alpha, cudaStream_t stream);
/* Split-K form for the weight gradients dW = dy^T a of the Linear layers (torch autograd of F.linear,
 * lightglue.py:139-148 etc.): one fp32 C [M,N] (row pitch ldc) = opA(A) * opB(B) over a very long K (the tokens).
 * The K range is cut into as many pieces as there are idle SMs; the fp32 partial tiles go to ws
 * (gemm_splitk_ws_floats(M,N,K) floats) and are summed in split order (deterministic).                 */
int64_t lgb200_gemm_splitk_ws_floats(int M, int N, int K);
int lgb200_gemm_bf16_splitk(const void* A, const void* B, float* C, int M, int N, int K, int a_mn_major,
                            int b_mn_major, int64_t lda, int64_t ldb, int64_t ldc, float* ws, cudaStream_t stream);
/* One nn.Linear-shaped GEMM on the same persistent tcgen05 kernel, with the epilogues the layer needs
 * (replaces F.linear / its autograd, lightglue.py:139-148, 156-163, 174-183, 195-221, 280):
 *   C[M,N] (c_dtype, row pitch ldc) = alpha * op(A) op(B)^T + bias[N]          (accumulate == 0)
 *   C[M,N] (fp32)                  += alpha * op(A) op(B)^T + bias[N]          (accumulate != 0; TMA reduce-add in L2)
 * A: [M,K] row pitch lda (or [K,M] when a_mn_major), B: [N,K] row pitch ldb (or [K,N] when b_mn_major), bf16;
 * bias: fp32 [N] or NULL.  forward: A = x, B = W;  dgrad: A = dy, B = W (b_mn_major);  row pitches let C be a column
 * block of a wider matrix (the [x | msg] FFN input, lightglue.py:162). */
int lgb200_linear(const void* A, const void* B, void* C, const float* bias, int M, int N, int K, int a_mn_major,
                  int b_mn_major, int64_t lda, int64_t ldb, int64_t ldc, int c_dtype, float alpha, int accumulate,
                  cudaStream_t stream);

/* ---- assignment head: sigmoid_log_double_softmax + NLL terms + argmax --------------------------------
 * replaces lightglue.py:256-268 (two log_softmax, transposed copy, slice assignment),
 * losses.py:6-25,62-73 (dense weight matrix and product) and the max() calls of
 * lightglue.py:86-90, 295-296.  sim [B,M,N] fp32.                                                 */
size_t lgb200_assign_ws_bytes(int B, int M, int N);
/* pass 1: lse_row [B,M] = logsumexp_j sim, lse_col [B,N] = logsumexp_i sim                         */
int lgb200_assign_lse(const float* sim, float* lse_row, float* lse_col, void* ws, int B, int M, int N,
                      cudaStream_t stream);
/* pass 2: scores_ij = (sim-lse_row_i) + (sim-lse_col_j) + (ls0_i + ls1_j) for i<M, j<N;
 *   scores [B,M+1,N+1] (optional dense output incl. dustbin column dust0, dustbin row dust1, corner 0)
 *   rowmax/rowarg [B,M]: max_j / argmax_j over j<N (lowest index wins ties); colmax/colarg [B,N] likewise
 *   pos_row_sum [B,M] (optional, needs gt [B,M,N] uint8): sum_j gt_ij (2 sim_ij - lse_row_i - lse_col_j)
 *   row_expsum [B,M] (optional): sum_{j<=N} exp(scores_ij)  (row_norm monitor, lightglue.py:596)     */
int lgb200_assign_scores(const float* sim, const float* lse_row, const float* lse_col, const float* ls0,
                         const float* ls1, const float* dust0, const float* dust1, const uint8_t* gt, float* scores,
                         float* rowmax, int* rowarg, float* colmax, int* colarg, float* pos_row_sum,
                         float* row_expsum, void* ws, int B, int M, int N, cudaStream_t stream);
/* backward of sum_ij gt_ij (2 sim_ij - lse_row_i - lse_col_j) scaled by gcoef[b]
 * (rowcnt_i = sum_j gt_ij, colcnt_j = sum_i gt_ij, as floats):
 *   dsim_ij = gcoef_b (2 gt_ij - exp(sim_ij-lse_row_i) rowcnt_i - exp(sim_ij-lse_col_j) colcnt_j)     */
int lgb200_assign_bwd(const float* sim, const float* lse_row, const float* lse_col, const uint8_t* gt,
                      const float* gcoef, const float* rowcnt, const float* colcnt, void* dsim, int out_dtype, int B,
                      int M, int N, cudaStream_t stream);
/* filter_matches (lightglue.py:293-309): mutual check + threshold; m0,m1 int64, -1 = no match       */
/* ---- assignment head fused with its similarity GEMM (csrc/assign_tc.cu; bf16 mdesc, D % 64 == 0, D <= 256) ----------
 * sim = alpha * mdesc0 . mdesc1^T (lightglue.py:283) is produced tile by tile in tensor memory and consumed there; it
 * is never written to memory.  md0: [B,M,D], md1: [B,N,D] bf16 contiguous.  Same outputs as lgb200_assign_lse /
 * lgb200_assign_scores (without the dense scores) / lgb200_assign_bwd followed by the two d(mdesc) GEMMs. */
int lgb200_assign_fused_lse(const void* md0, const void* md1, float alpha, float* lse_row, float* lse_col, int B, int M,
                            int N, int D, cudaStream_t stream);
/* gt ([B,M,N] bool) and pos_row_sum ([B,M]) are both set or both NULL. */
int lgb200_assign_fused_stats(const void* md0, const void* md1, float alpha, const float* lse_row, const float* lse_col,
                              const float* ls0, const float* ls1, const uint8_t* gt, float* rowmax, int* rowarg,
                              float* colmax, int* colarg, float* pos_row_sum, int B, int M, int N, int D,
                              cudaStream_t stream);
/* dmd0 [B*M, D], dmd1 [B*N, D] bf16 = dsim . md1, dsim^T . md0 with
 * dsim = gcoef[b] (2 gt - softmax_row rowcnt - softmax_col colcnt); gt_t = the [B,N,M] transpose of gt. */
int lgb200_assign_fused_bwd(const void* md0, const void* md1, float alpha, const float* lse_row, const float* lse_col,
                            const uint8_t* gt, const uint8_t* gt_t, const float* gcoef, const float* rowcnt,
                            const float* colcnt, void* dmd0, void* dmd1, int B, int M, int N, int D,
                            cudaStream_t stream);
int lgb200_filter_matches(const float* rowmax, const int* rowarg, const int* colarg, float th, int64_t* m0,
                          int64_t* m1, float* ms0, float* ms1, int B, int M, int N, cudaStream_t stream);

/* O(M+N) terms of one supervised layer (replaces the logsigmoid / weighted-sum / clamp / BCE tensor ops of
 * lightglue.py:264-267, losses.py:13-25, lightglue.py:81-94).  zt [T,2] = per token (matchability logit,
 * token-confidence logit), tokens ordered [image0 (B*M); image1 (B*N)].
 *   head_logsig   : ls = log sigmoid(z), du = log sigmoid(-z)                      [T] each
 *   head_terms_fwd: nll_pos = -(sum pos_row_sum + sum rowcnt ls0 + sum colcnt ls1)/num_pos,
 *                   nll_neg = -(sum neg0 du0 + sum neg1 du1)/num_neg, nll = bal nll_pos + (1-bal) nll_neg,
 *                   conf = token-confidence BCE against [argmax incl. dustbin == fin] (0 when fin* NULL)
 *   head_terms_bwd: d zt for upstream g_nll[B], g_conf[B]                                           */
int lgb200_head_logsig(const float* zt, float* ls, float* du, int64_t T, cudaStream_t stream);
/* The two per-token linear heads in one pass over x [T,D] fp32 (D % 8 == 0, D <= 512):
 *   head_token_fwd: zt[t] = (x[t].wm + bm, x[t].wt + bt)  (wt/bt NULL: column 1 repeats column 0),
 *                   ls/du as head_logsig, x_cast (nullable) = x in `dtype` for final_proj.
 *   head_token_bwd: dx = float(dmdw) + dzt[:,0] wm   (dmdw [T,D] in `dtype` = d md . W_final_proj; the confidence
 *                   head reads a detached x, lightglue.py:82-83), dW2 [2,D] = dzt^T x, db2 [2] = column sums of dzt.
 *                   ws: head_token_bwd_ws_floats(D) floats; counter: one uint32, zero before the first call. */
int lgb200_head_token_bwd_ws_floats(int D);
int lgb200_head_token_fwd(const float* x, const float* wm, const float* bm, const float* wt, const float* bt,
                          void* x_cast, float* zt, float* ls, float* du, int64_t ntok, int D, int dtype,
                          cudaStream_t stream);
int lgb200_head_token_bwd(const float* x, const void* dmdw, const float* dzt, const float* wm, float* dx, float* dW2,
                          float* db2, float* ws, unsigned* counter, int64_t ntok, int D, int dtype,
                          cudaStream_t stream);
int lgb200_head_terms_fwd(const float* zt, const float* pos_row_sum, const float* rowcnt, const float* colcnt,
                          const float* neg0, const float* neg1, const float* rowmax, const int* rowarg,
                          const float* colmax, const int* colarg, const int* fin0, const int* fin1,
                          const float* num_pos, const float* num_neg, float bal, float* nll, float* nll_pos,
                          float* nll_neg, float* conf, float* ws /* 4*B*ceil((M+N)/256) floats, 16-byte aligned */,
                          unsigned* counters /* B uint32, zero before the first call, left zero */, int B, int M, int N,
                          cudaStream_t stream);
int lgb200_head_terms_bwd(const float* zt, const float* rowcnt, const float* colcnt, const float* neg0,
                          const float* neg1, const float* rowmax, const int* rowarg, const float* colmax,
                          const int* colarg, const int* fin0, const int* fin1, const float* num_pos,
                          const float* num_neg, float bal, const float* g_nll, const float* g_conf, float* dzt, int B,
                          int M, int N, cudaStream_t stream);

/* ---- other assignment heads on the path ----------------------------------------------------------
 * log_double_softmax with a learned bin (gluestick.py:772-783) and log-domain Sinkhorn optimal
 * transport (gluefactory_nonfree/superglue.py:186-214).  scores/out [B,M+1,N+1].
 * ws: lgb200_heads_ws_bytes(B, M, N) bytes of scratch.                                               */
size_t lgb200_heads_ws_bytes(int B, int M, int N);
int lgb200_log_double_softmax(const float* sim, float bin_score, float* scores, void* ws, int B, int M, int N,
                              cudaStream_t stream);
/* the same with the (learnt) bin score read from device memory: no host read-back, capturable into a CUDA graph */
int lgb200_log_double_softmax_dev(const float* sim, const float* bin_score_dev, float* scores, void* ws, int B, int M,
                                  int N, cudaStream_t stream);
int lgb200_sinkhorn(const float* sim, float alpha, int iters, float* out, void* ws, int B, int M, int N,
                    cudaStream_t stream);
/* The same iterations with their potentials kept (uh [iters,B,M+1], vh [iters,B,N+1], both or neither; out may be
 * NULL), and the reverse sweep that turns grad = dL/dout [B,M+1,N+1] into dsim [B,M,N] plus the gradient entering the
 * dustbin row dzr [B,N+1] and column dzc [B,M] (d alpha = sum of both): the autograd of superglue.py:186-214
 * without its tape.  One persistent cooperative launch each; iters > 0 for the backward.                     */
int lgb200_sinkhorn_fwd(const float* sim, float alpha, int iters, float* out, float* uh, float* vh, void* ws, int B,
                        int M, int N, cudaStream_t stream);
int lgb200_sinkhorn_bwd(const float* sim, float alpha, int iters, const float* grad, const float* uh, const float* vh,
                        float* dsim, float* dzr, float* dzc, void* ws, int B, int M, int N, cudaStream_t stream);

/* column sums of a tall [rows, cols] matrix = bias gradient of an nn.Linear (autograd of lightglue.py:156 etc.).
 * ws: lgb200_colsum_slabs(rows, cols) * cols floats of scratch; counters: (cols+63)/64 uint32, zero before the
 * FIRST call (the kernel leaves them zero again); cols % 8 == 0.  One launch, deterministic summation order. */
int lgb200_colsum_slabs(int64_t rows, int cols);
/* rowcnt [B,M] = sum_j mask, colcnt [B,N] = sum_i mask of a 0/1 byte mask [B,M,N] (the `gt.sum(2)` / `gt.sum(1)` of
 * lightglue.py:595-600) in one pass; N % 16 == 0, N <= 4096, mask 16-byte aligned; counts are exact.            */
/* Weight gradient of the position encoder's projection (autograd of lightglue.py:37-44, theta = kp Wr^T):
 * part [lgb200_posenc_wgrad_blocks()][C][KD] per-CTA partials of sum_t g[t][c] kp[t][d]; g [T,C] fp32 with C == 32,
 * kp [T,KD] fp32, KD <= 4.  The caller sums the partials over their first dimension.                          */
int lgb200_posenc_wgrad_blocks(void);
int lgb200_posenc_wgrad(const float* g, const float* kp, float* part, int64_t T, int C, int KD, cudaStream_t stream);
int lgb200_mask_counts(const uint8_t* mask, float* rowcnt, float* colcnt, int B, int M, int N, cudaStream_t stream);
int lgb200_colsum(const void* a, float* out, float* ws, unsigned* counters, int64_t rows, int cols, int dtype,
                  cudaStream_t stream);

/* ---- flat-buffer optimiser (train.py:358-361, 513) and casts --------------------------------------- */
/* Adam over the flat buffers, one launch.  g is multiplied by grad_scale (1/world for the summed all-reduce).
 * step: 1-based step count for the bias correction; the four *_dev pointers (each may be NULL) move the per-step
 * controls into device memory so that ONE captured CUDA graph of the whole training step stays valid while they
 * change:  step_dev (count), lr_dev (scheduler, train.py:347-367), loss_scale_dev (GradScaler scale: gradients are
 * additionally multiplied by 1 / *loss_scale_dev, train.py:490), found_inf_dev (!= 0: the update is skipped --
 * GradScaler.step on overflow and the NaN guard of train.py:477-480, 503-512).
 * lr_scale_per_elem (may be NULL): per-element LR multiplier = the reference's lr_scaling groups (train.py:353-361). */
int lgb200_adam_flat(float* p, const float* g, float* m, float* v, int64_t n, const float* lr_scale_per_elem, float lr,
                     float beta1, float beta2, float eps, float weight_decay, int step, const int* step_dev,
                     float grad_scale, const float* lr_dev, const float* loss_scale_dev, const float* found_inf_dev,
                     cudaStream_t stream);
/* found_inf[0] = 1.0f if any of g[0..n) is non-finite, else 0.0f (GradScaler.unscale_'s check, no host sync). */
int lgb200_flat_grad_check(const float* g, int64_t n, float* found_inf, cudaStream_t stream);
/* One step of bookkeeping on the device: ++*step_dev unless *found_inf, and the GradScaler.update rule on
 * *loss_scale / *growth_tracker (both may be NULL): x backoff_factor on overflow, x growth_factor after
 * growth_interval consecutive clean steps. */
int lgb200_amp_update(const float* found_inf, int* step_dev, float* loss_scale, int* growth_tracker,
                      float growth_factor, float backoff_factor, int growth_interval, cudaStream_t stream);
int lgb200_cast_bf16(const float* src, void* dst, int64_t n, cudaStream_t stream);
/* x_out = x + y for two fp32 streams (x_out may alias x) and the compute-dtype copy of the sum: the merge of the two
 * gradients that reach a layer's output (from the next layer and from the layer's supervision head) fused with the cast
 * the layer's backward starts with -- replaces autograd's accumulation kernel + a cast.  n elements, dtype of x_cast. */
int lgb200_add_f32_cast(const float* x, const float* y, float* x_out, void* x_cast, int64_t n, int dtype,
                        cudaStream_t stream);
/* residual update of the fp32 stream fused with the cast for the next GEMM (lightglue.py:163, 219-220):
 * x_out = x + y (y in `dtype`, may be NULL), x_cast = (dtype) x_out; either output may be NULL.   */
int lgb200_residual_add_cast(const float* x, const void* y, float* x_out, void* x_cast, int64_t n, int dtype,
                             cudaStream_t stream);
/* same over [rows, cols] with the cast copy written at row pitch cast_pitch >= cols (elements): the copy lands
 * directly in the left / right half of the [x | msg] operand of ffn.0 (lightglue.py:162, 219), no torch.cat.   */
int lgb200_residual_add_cast_pitched(const float* x, const void* y, float* x_out, void* x_cast, int64_t rows,
                                     int64_t cols, int64_t cast_pitch, int dtype, cudaStream_t stream);

/* ---- ground-truth correspondences from a homography ---------------------------------
 * replaces the O(M N) part of gt_matches_from_homography (geometry/gt_generation.py:109-161): kp0 [B,M,2],
 * kp1 [B,N,2] pixel coordinates, kp0_1 = warp(kp0, H), kp1_0 = warp(kp1, H^-1) (the O(M+N) warps of
 * homography.py:161-180 are done by the caller).  dist = max(|kp0_1 - kp1|^2, |kp0 - kp1_0|^2); mutual nearest
 * neighbours with dist < pos_th^2 are positives; matches: index, -1 (best one-way error > neg_th^2), -2 (ignored).
 * assignment: [B,M,N] bytes (bool), cleared and filled by the call, or NULL.                              */
size_t lgb200_gt_homography_ws_bytes(int B, int M, int N);
int lgb200_gt_from_homography(const float* kp0, const float* kp1, const float* kp0_1, const float* kp1_0, float pos_th,
                              float neg_th, int64_t* matches0, int64_t* matches1, uint8_t* assignment, void* ws, int B,
                              int M, int N, cudaStream_t stream);
/* Same O(M N) label pass for ANY pair of reprojections (pose + depth ground truth, geometry/gt_generation.py:13-106):
 * vis0 [B,M] / vis1 [B,N] (bool, both or neither): the joint distance of a pair counts only when both points are
 * visible in the other view; valid0 / valid1 (bool, both or neither): "unmatched" (-1) additionally needs a valid depth.
 * assignment_t (may be NULL): the [B,N,M] transpose of `assignment`, written in the same pass (the fused assignment
 * backward walks the mask by columns too).  ws: lgb200_gt_homography_ws_bytes(B, M, N).  With all four masks NULL this is
 * lgb200_gt_from_homography. */
int lgb200_gt_from_reprojection(const float* kp0, const float* kp1, const float* kp0_1, const float* kp1_0,
                                const uint8_t* vis0, const uint8_t* vis1, const uint8_t* valid0, const uint8_t* valid1,
                                float pos_th, float neg_th, int64_t* m0, int64_t* m1, uint8_t* assignment,
                                uint8_t* assignment_t, void* ws, int B, int M, int N, cudaStream_t stream);
/* Extra unmatched labels from epipolar geometry (gt_generation.py:82-90; th_epi of matchers/depth_matcher.py): points
 * without valid depth that are still "ignore" (-2) become -1 when every still-ignored point of the other view is
 * further than th from their epipolar line (symmetric distance of geometry/epipolar.py:59-72, F [B,3,3] row-major).
 * m0 / m1 are updated in place; ws: B * (M + N) bytes. */
int lgb200_gt_epipolar_unmatched(const float* kp0, const float* kp1, const float* F, const uint8_t* valid0,
                                 const uint8_t* valid1, float th, int64_t* m0, int64_t* m1, uint8_t* ws, int B, int M,
                                 int N, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LGB200_H_ */
