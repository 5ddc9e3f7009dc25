/* liblafs_b200 -- C ABI of the B200-native LAFS hot path.
 *
 * The reference (szlbiubiubiu/LAFS_CVPR2024) is pure Python/PyTorch and has no FFI of its
 * own; the drop-in boundary is its Python module surface.  The Python
 * classes in lafs_cvpr2024_b200/ keep that surface and bind the entry points below with
 * ctypes.  Each entry point cites the reference code it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host;
 *   - the caller owns every buffer; the library never allocates device memory and keeps
 *     no mutable global state except a thread-local last-error string;
 *   - work is enqueued on `stream` (a cudaStream_t) and the call returns immediately;
 *   - return value: 0 on success, negative on failure (LAFS_ERR_*); the message is
 *     available from lafs_last_error_string().  No exception crosses the boundary;
 *   - dtype codes: 0 = fp32, 1 = bf16, 2 = fp16.
 */
#ifndef LAFS_B200_H_
#define LAFS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* lafs_stream_t; /* cudaStream_t */

#define LAFS_API __attribute__((visibility("default")))

#define LAFS_OK 0
#define LAFS_ERR_ARG (-1)
#define LAFS_ERR_CUDA (-2)
#define LAFS_ERR_WORKSPACE (-3)

#define LAFS_F32 0
#define LAFS_BF16 1
#define LAFS_F16 2
#define LAFS_U8 3 /* images only: lafs_gather_embed_fwd */

LAFS_API int lafs_version(void);
LAFS_API const char* lafs_last_error_string(void);
/* 1 when the running device is compute capability 10.x, 0 otherwise, negative on error. */
LAFS_API int lafs_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * (3) Teacher EMA  --  replaces the per-parameter loop  lafs_train.py:610-613
 *       param_k.data.mul_(m).add_((1 - m) * param_q.detach().data)
 * One launch for all tensors.  `table` is a device array of nchunks records
 *   struct { float* k; const float* q; int64_t n; }   (24 bytes, n <= LAFS_EMA_CHUNK)
 * each describing a contiguous run of one fp32 tensor pair.  Arithmetic is
 *   k = fl(fl(k*m) + fl(q*one_minus_m))   -- three separately rounded fp32 operations,
 * bit-identical to the reference (m and 1-m are rounded to fp32 by the caller from the
 * float64 schedule value, as PyTorch does).  max_ctas <= 0: one CTA per chunk (fastest when the
 * kernel runs alone); max_ctas > 0: that many persistent CTAs (e.g. 148 = one per SM) so that the
 * kernel can run concurrently with another one on a second stream.
 */
#define LAFS_EMA_CHUNK 16384
LAFS_API int lafs_ema_multi(const void* table, int nchunks, float m, float one_minus_m, int max_ctas,
                            lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (2) DINO loss  --  replaces DINOLoss.forward / update_center  lafs_train.py:643-679
 *
 * student [ncrops*B, K] (crop v occupies rows [v*B,(v+1)*B): Tensor.chunk(ncrops)),
 * teacher [2*B, K], both of `dtype`; center [K] fp32 (the OLD centre: the update happens
 * after the loss).  inv_student_temp = 1/student_temp, inv_teacher_temp =
 * 1/teacher_temp_schedule[epoch].
 *
 * lafs_dino_fwd writes
 *   loss_out   [1]  fp32   mean over the 2*ncrops-2 (teacher view, student crop) terms
 *   row_stats  [(ncrops+2)*B] fp32: log2-domain log-sum-exp of every student row, followed
 *              by the 2*B teacher rows (consumed by lafs_dino_bwd)
 *   colsum_out [K]  fp32   sum over the 2*B teacher rows (the message of the reference's
 *              dist.all_reduce, lafs_train.py:674-675)
 * center_out (may be NULL): single-process shortcut -- also writes the updated centre
 *   center*momentum + (colsum/(2B))*(1-momentum) in the same pass (no all-reduce needed).
 * The result is deterministic (no floating-point atomics).
 * K must be a multiple of 8 (16-bit dtypes) or 4 (fp32); 2 <= ncrops <= 12.
 */
LAFS_API size_t lafs_dino_workspace_bytes(int B, int K, int ncrops);
LAFS_API int lafs_dino_fwd(const void* student, const void* teacher, const float* center, int B, int K,
                  int ncrops, float inv_student_temp, float inv_teacher_temp, int dtype,
                  float* loss_out, float* row_stats, float* colsum_out, void* workspace,
                  size_t workspace_bytes, float* center_out, float momentum, float one_minus_momentum,
                  lafs_stream_t stream);
/* grad_student[v*B+b, k] = grad_out * d loss / d student; grad_out is a device scalar
 * (the upstream gradient; autograd hands it over on the device). */
LAFS_API int lafs_dino_bwd(const void* student, const void* teacher, const float* center,
                  const float* row_stats, const float* grad_out, int B, int K, int ncrops,
                  float inv_student_temp, float inv_teacher_temp, int dtype, void* grad_student,
                  lafs_stream_t stream);
/* Forward and backward from one call (no autograd round trip), for training loops where the upstream
 * gradient of the loss is known when the loss is computed (grad_out: device scalar, 1 for a root
 * loss).  Can process the batch in waves of samples (env LAFS_DINO_WAVE) so that the gradient pass
 * re-reads from L2; measured slower than one wave on B200, which is the default.  Outputs as
 * lafs_dino_fwd + lafs_dino_bwd; workspace from lafs_dino_fused_workspace_bytes. */
LAFS_API size_t lafs_dino_fused_workspace_bytes(int B, int K, int ncrops);
LAFS_API int lafs_dino_fwd_bwd(const void* student, const void* teacher, const float* center,
                               const float* grad_out, int B, int K, int ncrops, float inv_student_temp,
                               float inv_teacher_temp, int dtype, float* loss_out, float* row_stats,
                               float* colsum_out, void* grad_student, void* workspace, size_t workspace_bytes,
                               float* center_out, float momentum, float one_minus_momentum,
                               lafs_stream_t stream);
/* center_out = center*momentum + (colsum/count)*(1-momentum)   lafs_train.py:676-679
 * (count = 2B*world_size; colsum already all-reduced by the caller when world>1).
 * center_out may alias center; the reference re-binds a NEW tensor, and the
 * backward pass needs the old centre, so the Python module passes a fresh buffer. */
LAFS_API int lafs_center_ema(const float* center, const float* colsum, float count, float momentum,
                    float one_minus_momentum, int K, float* center_out, lafs_stream_t stream);
/* Column sum over `rows` rows of x [rows,K] (torch.sum(teacher_output, dim=0),
 * lafs_train.py:674) for update_center called on its own; workspace >= 32*K*4 bytes. */
LAFS_API int lafs_colsum(const void* x, int rows, int K, int dtype, float* out, void* workspace,
                size_t workspace_bytes, lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (1) Landmark post-processing and patch gather
 *
 * lafs_landmark_post replaces face_pre_pro/ViT_face.py:1347-1378 (and :694-705):
 *   theta = (raw - min)/(max - min)*scale  jointly over the 2n numbers of a sample,
 *   view [n,2], + noise (may be NULL), gather rows extract_id (may be NULL; int64 [B,keep]).
 * theta_out [B, keep or n, 2] fp32.  minmax_out (may be NULL) [B,2] receives (min,max).
 */
LAFS_API int lafs_landmark_post(const float* raw, const float* noise, const int64_t* extract_id,
                       float* theta_out, float* minmax_out, int B, int n, int keep, float scale,
                       lafs_stream_t stream);
/* backward of the plain (no gather) form: grad_raw [B,2n] from grad_theta [B,n,2]. */
LAFS_API int lafs_landmark_post_bwd(const float* raw, const float* grad_theta, float* grad_raw, int B,
                           int n, float scale, lafs_stream_t stream);

/* lafs_gather_fwd replaces extract_patches_pytorch_gridsample,
 * face_pre_pro/ViT_face.py:1615-1656 (+ the einops rearrange lafs_train.py:538 when
 * layout = LAFS_LAYOUT_TOKENS).  imgs [Bv,C,H,W] fp32, theta [Bv,n,2] fp32 (x,y) pixels,
 * 8x8 patches, bilinear, zero padding, align_corners=False.
 *   LAFS_LAYOUT_MOSAIC : out [Bv,C,8r,8r], r*r = n   (the reference's return value)
 *   LAFS_LAYOUT_TOKENS : out [Bv,n,64*C], feature (i*8+j)*C+c
 * coord_mode LAFS_COORD_DIV reproduces the reference's CPU arithmetic bit for bit
 * (grid = (offset+theta)/(H/2) - 1 with an IEEE division); LAFS_COORD_RECIP reproduces
 * eager CUDA, which multiplies by fp32(2/H).
 */
#define LAFS_LAYOUT_MOSAIC 0
#define LAFS_LAYOUT_TOKENS 1
#define LAFS_COORD_DIV 0
#define LAFS_COORD_RECIP 1
LAFS_API int lafs_gather_fwd(const float* imgs, const float* theta, float* out, int Bv, int C, int H, int W,
                    int n, int layout, int coord_mode, lafs_stream_t stream);
/* grad_out in `layout`; grad_imgs (may be NULL) [Bv,C,H,W] must be zero-filled by the
 * caller (scatter-add); grad_theta (may be NULL) [Bv,n,2]. */
LAFS_API int lafs_gather_bwd(const float* imgs, const float* theta, const float* grad_out, float* grad_imgs,
                    float* grad_theta, int Bv, int C, int H, int W, int n, int layout,
                    int coord_mode, lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (4) CosFace / ArcFace margin head with fused softmax cross-entropy
 *     replaces CosFace.forward (face_pre_pro/ViT_face.py:49-89) + the loss applied to its
 *     output (train_largescale.py:601-604,820).  ArcFace is named by the reference
 *     (ViT_face.py:416-417,654-655) but never defined there: paper form, parity unpinned.
 *
 * Operands are bf16 and L2-normalised per row: lafs_normalize_rows(x [R,D] of dtype) writes
 * x/max(||x||,1e-12) as bf16 (F.normalize, ViT_face.py:53) and optionally 1/max(||x||,eps).
 * Classes are sharded with torch.chunk's rule (ViT_face.py:56): a rank owns global classes
 * [class_lo, class_lo + C_local).  label_a [B] int64 global ids; label_b (may be NULL) is the
 * mixup partner (util/mixup_my.py:18-24) with weight 1-lam; kind 0 = CosFace s*(cos - m*t),
 * 1 = ArcFace (hard labels).  D must be a multiple of 64, D <= 768.
 *
 * lafs_head_fwd   -> row_stats [B,4] = (max, sum-exp, z[label_a], z[label_b]) over the LOCAL
 *                    classes, max in the log2 domain; no [B,C] tensor is written.
 * lafs_head_merge -> merges nparts records [nparts,B,4] (one per rank) into [B,4].
 * lafs_head_loss  -> row_lse2 [B] (log2-domain lse, kept for backward) and the mean loss.
 * lafs_head_logits-> the full fp32 logits [B, C_local] (row stride ldc) for API parity with
 *                    CosFace.forward.
 * lafs_head_grad_logits -> G = (softmax - target) * d z/d cos * gscale * (*grad_out) as bf16
 *                    [B, C_local] (row stride ldg, a multiple of 8), recomputing the logits on the
 *                    tensor cores; gscale = s / B_global, grad_out = device scalar upstream gradient.
 * lafs_head_bwd_embed  -> grad_e_hat [B,D] fp32 = G . W_hat (split-K GEMM; sharded heads all-reduce
 *                    it over ranks before the Jacobian).  workspace: lafs_head_bwd_workspace_bytes.
 * lafs_head_bwd_weight -> grad_w [C_local,D] fp32 = normalize-backward(G^T . E_hat).
 * lafs_normalize_bwd   -> out = (g - x_hat <x_hat, g>) * inv_norm per row (F.normalize backward).
 */
LAFS_API int lafs_normalize_rows(const void* x, int dtype, int R, int D, void* out_bf16, float* inv_norm,
                                 lafs_stream_t stream);
LAFS_API size_t lafs_head_workspace_bytes(int B, int C_local, int D);
LAFS_API int lafs_head_fwd(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                           float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                           float* row_stats, void* workspace, size_t workspace_bytes, lafs_stream_t stream);
LAFS_API int lafs_head_merge(const float* parts, int nparts, int B, float* row_stats, lafs_stream_t stream);
LAFS_API int lafs_head_loss(const float* row_stats, const int64_t* label_a, const int64_t* label_b, float lam, int B,
                            float* row_lse2, float* loss_out, lafs_stream_t stream);
LAFS_API int lafs_head_logits(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                              float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                              float* logits, long long ldc, lafs_stream_t stream);
LAFS_API int lafs_head_grad_logits(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                                   float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                                   const float* row_lse2, const float* grad_out, float gscale, void* grad_bf16,
                                   long long ldg, lafs_stream_t stream);
LAFS_API size_t lafs_head_bwd_workspace_bytes(int B, int C_local, int D);
LAFS_API int lafs_head_bwd_embed(const void* grad_bf16, long long ldg, const void* w_hat, int B, int C_local, int D,
                                 float* grad_e_hat, void* workspace, size_t workspace_bytes, lafs_stream_t stream);
LAFS_API int lafs_head_bwd_weight(const void* grad_bf16, long long ldg, const void* e_hat, const void* w_hat,
                                  const float* inv_norm_w, int B, int C_local, int D, float* grad_w,
                                  lafs_stream_t stream);
LAFS_API int lafs_normalize_bwd(const float* g, const void* x_hat_bf16, const float* inv_norm, int R, int D, float* out,
                                lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (1b) Fused landmark gather -> patch embedding on tensor cores
 *      replaces extract_patches_pytorch_gridsample + rearrange + patch_to_embedding
 *      (ViT_face.py:1615-1656, lafs_train.py:538, ViT_face.py:760-761) for 112x112 faces.
 * lafs_embed_weight_prep: nn.Linear(192, dim).weight fp32 [dim,192] (feature (i*8+j)*3+c)
 *      -> bf16 [dim,192] in the kernel's K order (c*64 + j*8 + i); bias (may be NULL = 0) is
 *      copied to bias_out (may be NULL) in the same launch.  Stack two models by
 *      writing them back to back ([2*dim,192]) to project the same patches with student and
 *      teacher weights (the two global SSL views, lafs_train.py:577-581).
 * lafs_gather_embed_fwd: out_m [Bv, n, dim] (out_dtype bf16, or fp32 for verification) =
 *      tokens(imgs, theta) @ W_m^T + bias_m for m < n_models; bias [n_models*dim] fp32;
 *      n <= 208, dim % 128 == 0.  Tokens and weights are rounded to bf16, accumulation is fp32.
 *      imgs is fp32 (in_dtype LAFS_F32: the reference's normalised tensors) or uint8 (LAFS_U8:
 *      the decoded pixels; the reference's ToTensor + Normalize (lafs_train.py:800-803) is applied
 *      in the kernel as in_scale*u8 + in_shift, e.g. 2/255 and -1, and the zero padding of
 *      grid_sample stays a zero in normalised space).
 */
LAFS_API int lafs_embed_weight_prep(const float* weight, const float* bias, int dim, void* out_bf16, float* bias_out,
                                    lafs_stream_t stream);
LAFS_API int lafs_gather_embed_fwd(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                   const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                   int Bv, int H, int W, int n, int dim, int n_models, lafs_stream_t stream);
/* Training form: additionally keeps the gathered tokens for the weight gradient, bf16 [Bv*n, 208] in the kernel's K
 * order (k = c*64 + j*8 + i); the kernel writes columns 0..191 only -- the caller sets column 192 to 1 and columns
 * 193..207 to 0 ONCE (the ones column yields the bias gradient in the same GEMM).  tokens_perm_out may be NULL. */
LAFS_API int lafs_gather_embed_fwd_save(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                        const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                        int Bv, int H, int W, int n, int dim, int n_models, void* tokens_perm_out,
                                        lafs_stream_t stream);

/* Sequence form: the kernel's epilogue also does what ViT_face.py:762-768 does after
 * patch_to_embedding -- `torch.cat((cls_tokens, x), 1); x += pos_embedding[:, :n+1]; x = dropout(x)` -- so out_m is
 * the transformer input [Bv, n+1, dim]: row 0 = cls_m + pos_m[0], row 1+t = embedding[t] + pos_m[1+t].  pos_m points at
 * pos_embedding's rows ([>= n+1, dim] fp32), cls_m at cls_token ([dim] fp32); pos0 == NULL selects the plain form.
 * drop_p in [0,1): inverted dropout with a counter-based mask (a function of drop_seed, model and output element;
 * torch's Philox stream is not reproduced -- the reference is stochastic here).  tokens_perm_out as above. */
LAFS_API int lafs_gather_embed_seq_fwd(const void* imgs, int in_dtype, float in_scale, float in_shift, const float* theta,
                                       const void* w_perm_bf16, const float* bias, void* out0, void* out1, int out_dtype,
                                       int Bv, int H, int W, int n, int dim, int n_models, void* tokens_perm_out,
                                       const float* pos0, const float* cls0, const float* pos1, const float* cls1,
                                       float drop_p, unsigned int drop_seed, lafs_stream_t stream);

/* Backward of patch_to_embedding (the training path of the fused kernel; the reference gets it from
 * autograd through nn.Linear, ViT_face.py:760-761 / lafs_train.py:544): two tcgen05 GEMMs over the
 * M = faces*landmarks token rows.  grad_emb [M,dim], tokens [M,192] ('(p1 p2 c)' order), weight
 * [dim,192] are bf16; grad_w [dim,192] and grad_tokens [M,192] are fp32 (grad_tokens feeds
 * lafs_gather_bwd with LAFS_LAYOUT_TOKENS).  dim % 8 == 0. */
LAFS_API size_t lafs_embed_bwd_workspace_bytes(int M, int dim);
LAFS_API int lafs_embed_bwd_weight(const void* grad_emb_bf16, const void* tokens_bf16, int M, int dim, float* grad_w,
                                   void* workspace, size_t workspace_bytes, lafs_stream_t stream);
/* grad_w [dim,192] ('(p1 p2 c)' order) and grad_b [dim] (may be NULL) of patch_to_embedding from the tokens
 * lafs_gather_embed_fwd_save kept ([M,208], ones column included): ONE split-K tcgen05 GEMM grad_emb^T . tokens plus a
 * [dim,193] reduce / un-permute kernel -- no re-gather, no fp32 token tensor, no separate column-sum pass for the
 * bias (lafs_train.py:600 through ViT_face.py:761).  accumulate != 0 adds to grad_w / grad_b (several view groups). */
LAFS_API int lafs_embed_bwd_weight_perm(const void* grad_emb_bf16, const void* tokens_perm_bf16, int M, int dim, float* grad_w,
                                        float* grad_b, int accumulate, void* workspace, size_t workspace_bytes,
                                        lafs_stream_t stream);
LAFS_API int lafs_embed_bwd_tokens(const void* grad_emb_bf16, const void* weight_bf16, int M, int dim,
                                   float* grad_tokens, lafs_stream_t stream);

/* The F.normalize Jacobian of dW on the tensor core (the default dW path of the Python wrappers; verified on B200
 * in round 2, tests/test_gpu_head.py::test_dw_jacobian_on_tensor_core_variant and the bench-shape tests).
 * lafs_head_grad_logits_t = lafs_head_grad_logits + per-class partial dots tpart[(mtile*4+quarter)*ldt + c]
 * (sum over a 32-row group of G[b,c]*cos[b,c]; 4*ceil(B/128) rows, ldt a multiple of 32 >= C_local rounded up
 * to 32, 128-byte aligned).  lafs_head_bwd_weight_t sums the rows (in place, row 0) to t[c] = <w_hat_c, dW_hat_c>
 * and computes grad_w = inv_norm_w * (G^T.E_hat - diag(t).W_hat) in ONE GEMM (two extra k blocks per tile). */
LAFS_API int lafs_head_grad_logits_t(const void* e_hat, const void* w_hat, const int64_t* label_a, const int64_t* label_b,
                                     float lam, int B, int C_local, int D, int class_lo, float s, float m, int kind,
                                     const float* row_lse2, const float* grad_out, float gscale, void* grad_bf16,
                                     long long ldg, float* tpart, long long ldt, lafs_stream_t stream);
LAFS_API int lafs_head_bwd_weight_t(const void* grad_bf16, long long ldg, const void* e_hat, const void* w_hat,
                                    const float* inv_norm_w, float* tpart, int tparts, long long ldt, int B,
                                    int C_local, int D, float* grad_w, lafs_stream_t stream);

/* out [M, N] fp32 = A^T . B, A [Kr rows, M cols], B [Kr rows, N cols] bf16 row-major (leading dimensions in
 * elements, multiples of 8): the dW GEMM without a Jacobian pass.  Used by the fused DINO head below. */
LAFS_API int lafs_gemm_tn(const void* a_bf16, long long lda, const void* b_bf16, long long ldb, int Kr, int M, int N,
                          float* out, long long ldo, lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (f1) DINOHead.last_layer fused with DINOLoss  --  the [(ncrops+2)B, K] logits never exist in HBM.
 *      replaces  x = nn.functional.normalize(x, dim=-1, p=2); x = self.last_layer(x)
 *      (vision_transformer.py:284,296-300; last_layer = weight_norm(Linear(bottleneck, K, bias=False))) followed by
 *      DINOLoss.forward / update_center (lafs_train.py:643-679).
 * The contractions are the margin head's tcgen05 GEMMs (lafs_head_fwd = per-row online-softmax statistics,
 * lafs_head_grad_logits with labels -1 = bf16 softmax probabilities, lafs_head_bwd_embed = P . W, lafs_gemm_tn);
 * the entry points below are the streaming passes around them (csrc/dino_head.cu explains the algebra; the Python
 * side is lafs_cvpr2024_b200/dino_head.py).
 *   lafs_dh_extra_cols   K columns appended to the TEACHER operands (64): features [x_hat | 1 1 1 | 0..], prototypes
 *                        [w | -c_hi -c_mid -c_lo | 0..] (three-term bf16 split of the fp32 centre), so that the
 *                        accumulator holds <x_hat, w_k> - center[k].
 *   lafs_dh_prep_rows    x [R, D] (any dtype code) -> F.normalize rows, bf16 [R, ld], ld = D or D + extra;
 *                        inv_norm [R] = 1/max(||x||, 1e-12) (may be NULL).
 *   lafs_dh_xsum         xsum [D] = sum over the R rows of x_hat (fp32, fixed order).
 *   lafs_dh_prep_weight  weight_norm rows w_k = v_k * (g_k/||v_k||) (weight_g NULL: g = 1) -> bf16 [K, ld];
 *                        inv_norm [K] = 1/||v_k||; with xsum: colsum [K] = <w_k, xsum> = the column sums of the
 *                        teacher logits (torch.sum(teacher_output, dim=0), lafs_train.py:674) for lafs_center_ema.
 *   lafs_dh_lse2         merged row statistics [R, 4] of lafs_head_fwd -> log2-domain lse [R].
 *   lafs_dh_loss         loss = 1/((2 ncrops-2) B) sum_{iq<2, v != iq, i} [ lse(s_v,i/ts) - <U_iq,i, x_hat_v,i>/ts ],
 *                        U [2B, D] = Q . W_s (fp32), x_hat_s bf16 [ncrops B, D]; sample_loss [B] = per-sample sums
 *                        (scratch / diagnostic), added in a fixed order.
 *   lafs_dh_bwd_rows     student rows: dx = F.normalize backward of coef (cnt_v O - sum_{iq != v} U_iq), O = P_s . W_s;
 *                        all rows: y [(ncrops+2) B, D] bf16 = the B operand of the dW GEMM (cnt_v x_hat_s | -X~).
 *                        coef = inv_student_temp / ((2 ncrops-2) B) * grad_out[0].
 *   lafs_dh_wn_bwd       weight_norm backward from the raw dW [K, D]: grad_v = coef*grad_out[0] * g/||v|| (dW - v_hat
 *                        <v_hat, dW>), grad_g [K] = coef*grad_out[0] * <v_hat, dW> (may be NULL).  grad_v may alias dw_raw. */
LAFS_API int lafs_dh_extra_cols(void);
LAFS_API int lafs_dh_prep_rows(const void* x, int dtype, int R, int D, int ld, void* out_bf16, float* inv_norm,
                               lafs_stream_t stream);
LAFS_API int lafs_dh_xsum(const void* x_hat_bf16, int R, int D, int ld, float* xsum, lafs_stream_t stream);
LAFS_API int lafs_dh_prep_weight(const float* weight_v, const float* weight_g, const float* center, const float* xsum,
                                 int K, int D, int ld, void* out_bf16, float* inv_norm, float* colsum,
                                 lafs_stream_t stream);
LAFS_API int lafs_dh_lse2(const float* row_stats, int R, float* lse2, lafs_stream_t stream);
LAFS_API int lafs_dh_loss(const float* lse2_s, const float* U, const void* x_hat_s_bf16, int B, int ncrops, int D,
                          float inv_student_temp, float* sample_loss, float* loss_out, lafs_stream_t stream);
LAFS_API int lafs_dh_bwd_rows(const float* O, const float* U, const void* x_hat_s_bf16, const float* inv_norm_s,
                              const float* grad_out, int B, int ncrops, int D, float inv_student_temp, float* dx,
                              void* y_bf16, lafs_stream_t stream);
LAFS_API int lafs_dh_wn_bwd(const float* dw_raw, const float* weight_v, const float* weight_g, const float* inv_norm,
                            const float* grad_out, int K, int D, float coef, float* grad_v, float* grad_g,
                            lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * (4e) Exchange steps of the class-sharded head over NVLink peer memory (one kernel each, instead of
 *      NCCL all_gather / all_reduce).  The reference has no counterpart (it replicates the head,
 *      ViT_face.py:56 only chunks the weight for the matmul); the ownership rule is torch.chunk's.
 * Every rank maps one symmetric buffer of lafs_xchg_bytes(world, B, D, offsets) bytes (same layout on
 * all ranks, zero-filled once, e.g. torch.distributed._symmetric_memory); peer_base is a DEVICE array
 * of the `world` base addresses as seen from this rank.  offsets[0..4] = flags, statistics slots,
 * partial dE_hat [B,D] fp32 (all-reduce input), summed dE_hat [B,D] fp32 (output), error flag (uint32,
 * non-zero after a peer failed to arrive within the spin bound).  world <= 8, D % 4 == 0.
 * All ranks must issue the same sequence of lafs_xchg_* calls. */
LAFS_API size_t lafs_xchg_bytes(int world, int B, int D, size_t* offsets);
LAFS_API int lafs_xchg_stats(const void* peer_base, int rank, int world, int B, int D, const float* local_stats,
                             float* merged_stats, lafs_stream_t stream);
LAFS_API int lafs_xchg_allreduce(const void* peer_base, int rank, int world, int B, int D, lafs_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * (f3) Student update: per-tensor gradient clipping + AdamW + teacher EMA as one multi-tensor call
 *      replaces utils.clip_gradients (utils.py:132-141: a norm kernel, a .item() host sync and a mul_ per tensor),
 *      utils.cancel_gradients_last_layer (utils.py:144-149), torch.optim.AdamW.step (lafs_train.py:400,601-609)
 *      and the teacher EMA loop (lafs_train.py:610-613).
 * table: device array of nchunks records {float* p; const float* g; float* m; float* v; float* k; int32 n; int32 tensor}
 *      (48 bytes; runs of <= LAFS_EMA_CHUNK elements; g == NULL: the tensor has no gradient this step, only its EMA
 *      runs; k == NULL: no EMA).  first_chunk [ntensors+1] (int32): the chunks of tensor t are
 *      [first_chunk[t], first_chunk[t+1]).  regularized [ntensors] (uint8): weight decay applies (utils.py:662-673).
 * hyper: DEVICE array of 10 floats, refreshed by the host every step (so a captured graph can be replayed):
 *      {1 - lr*wd, lr/(1-b1^t), sqrt(1-b2^t), eps, 1-b1, b2, 1-b2, ema_m, 1-ema_m, clip (<= 0: off)}.
 * Outputs: grad_norms [ntensors] (what clip_gradients returns, left on the device), clip_coef [ntensors].
 * Arithmetic = torch.optim.AdamW's single-tensor formulas in fp32; the EMA uses the NEW parameter value. */
LAFS_API size_t lafs_optim_workspace_bytes(int nchunks, int ntensors);
LAFS_API int lafs_adamw_ema_multi(const void* table, int nchunks, const int* first_chunk, const unsigned char* regularized,
                                  int ntensors, const float* hyper, float* grad_norms, float* clip_coef, void* workspace,
                                  size_t workspace_bytes, lafs_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* LAFS_B200_H_ */
