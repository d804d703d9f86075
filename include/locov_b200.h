/*
 * locov_b200.h — C ABI of liblocov_b200.so: the B200 (sm_100a) implementation of LocOV's
 * region-text matching hot path.
 *
 * The reference (lmb-freiburg/locov) is pure Python on top of Detectron2 and exposes no FFI; its
 * "plugin API" for this path is the set of nn.Module classes registered in Detectron2's registries
 * (SURVEY.md §8b).  Each entry point below therefore cites the reference *call site* whose library
 * kernel(s) it replaces.  The Python host side (package locov_b200) binds these symbols with ctypes
 * and re-exposes them behind the reference's module interfaces.
 *
 * Conventions
 *   - every function returns int: 0 = LOCO_OK, <0 = LOCO_E_*, >0 = a cudaError_t value;
 *     loco_last_error() returns a thread-local human-readable message for the last failure.
 *   - all pointers are DEVICE pointers unless the parameter is called `host_*`; the caller owns all
 *     memory; the library never allocates user-visible memory and never synchronises `stream`.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - bf16 tensors are raw uint16 bit patterns; "ld" = leading dimension in ELEMENTS (row stride).
 *     Every bf16 matrix handed to a tensor-core entry point must have a 16-byte aligned base and
 *     ld % 8 == 0 (TMA global-stride rule), otherwise LOCO_E_ALIGN.
 *   - "hi/lo" operand pairs implement the fp32-accurate mode: x ≈ hi + lo with hi = bf16(x),
 *     lo = bf16(x - hi); a product is evaluated as hi*hi + hi*lo + lo*hi on the bf16 tensor pipe with
 *     fp32 accumulation in TMEM (≈2^-16 relative per product).  Passing lo == NULL selects plain
 *     bf16 mode.
 *   - thread-safe and re-entrant (autograd calls backward entries from another host thread).
 */
#ifndef LOCOV_B200_H
#define LOCOV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LOCO_OK 0
#define LOCO_E_BADARG (-1)
#define LOCO_E_UNSUPPORTED (-2)
#define LOCO_E_ALIGN (-3)
#define LOCO_E_DEVICE (-4)
#define LOCO_E_DRIVER (-5)

/* tensor element types */
#define LOCO_F32 0
#define LOCO_BF16 1
/* feature-map layouts */
#define LOCO_NCHW 0
#define LOCO_NHWC 1
/* LSM alignment modes — MODEL.MMSS_HEAD.GROUNDING.ALIGNMENT, reference grounding_head.py:162-175 */
#define LOCO_ALIGN_SOFTMAX 0
#define LOCO_ALIGN_HARDMAX 1

/* ---- library / device ------------------------------------------------------------------------- */
int loco_version(void);                 /* 10000*major + 100*minor + patch */
const char *loco_last_error(void);      /* thread-local; never NULL */
int loco_device_check(int device);      /* LOCO_OK iff `device` is compute capability 10.x */
int loco_sm_count(int device);          /* number of SMs (148 on B200) or <0 */
long long loco_launch_count(void);      /* kernels this library has launched so far (process-wide, monotonic) */
/* developer probe (env LOCOV_B200_TIMELINE=1): copies the per-CTA phase stamps (16 x uint64 nanoseconds per CTA: entry, setup
 * done, first operand stage landed, last MMA issued, last accumulator complete, epilogue done, exit, unused, then 8 policy-defined
 * stamps inside the epilogue) of the most recent
 * tensor-core launch to `host`; returns the number of CTAs copied (0 when the probe is off).  Synchronises the device. */
int loco_debug_timeline_read(unsigned long long *host, int max_ctas);

/* ---- RoIAlign ----------------------------------------------------------------------------------
 * Replaces: roi_emb_heads.py:182-187,243-245  self.pooler(features, boxes)
 *           -> Detectron2 ROIPooler/ROIAlign -> torchvision.ops.roi_align(aligned=True) CUDA kernel.
 * feat  [N,C,H,W] (LOCO_NCHW) or [N,H,W,C] (LOCO_NHWC), fp32.
 * rois  [R,5] fp32 rows (batch_idx, x1, y1, x2, y2) in image coordinates.
 * out   out_layout LOCO_NCHW: [R,C,PH,PW] fp32 (what the reference's res5 stage receives); LOCO_NHWC: [R,PH,PW,C] channels-last in
 *       fp32 or bf16 (out_dtype) — what cuDNN's tensor-core convolutions of res5 want (torch.channels_last), half the bytes of the
 *       dominant write of the path in bf16, and no shared-memory staging (a warp writes 512 / 256 contiguous bytes per bin).
 *       Channels-last output needs C % 4 == 0 and a 16-byte aligned pointer.
 * Sampling-grid coordinates and integer tap indices are bit-exact with the float32 reference
 * arithmetic (loco_roi_align_grid_dump exposes them); pooled values are toleranced (1e-4 rel).
 * workspace: loco_roi_align_workspace_bytes() bytes: an NCHW map is transposed to channels-last into it once per call, and the
 * launch order of the rois (heaviest first: a small bucket-sort launch that removes the tail of late, large rois) lives behind it.
 * NULL is accepted for LOCO_NHWC features (rois are then taken in the given order).
 */
int64_t loco_roi_align_workspace_bytes(int N, int C, int H, int W, int feat_layout, int R);
int loco_roi_align_fwd(const float *feat, int N, int C, int H, int W, int feat_layout,
                       const float *rois, int R, int PH, int PW, float spatial_scale,
                       int sampling_ratio, int aligned, void *out, int out_layout, int out_dtype, void *workspace,
                       void *stream);

/* Backward of the above (torchvision roi_align_backward semantics): dfeat [N,C,H,W] fp32 = gradient w.r.t. the feature
 * map.  With `workspace` (loco_roi_align_bwd_workspace_bytes() bytes, 16-byte aligned, contents ignored) and C % 4 == 0 the
 * gradient is accumulated channels-last with 128-bit vector atomics and transposed into dfeat, which is then fully
 * overwritten; without it (NULL) contributions are added to dfeat with scalar atomics and dfeat must be zero-filled by
 * the caller.  Passing a zero-filled dfeat is correct in both cases. */
int64_t loco_roi_align_bwd_workspace_bytes(int N, int C, int H, int W, int R);
int loco_roi_align_bwd(const float *dout, int N, int C, int H, int W, const float *rois, int R,
                       int PH, int PW, float spatial_scale, int sampling_ratio, int aligned,
                       float *dfeat, void *workspace, void *stream);


/* Spatial mean of the res5 output: replaces ``box_features.mean(dim=[2, 3])`` (reference ovr/modeling/roi_heads/roi_emb_heads.py:262,
 * :329 and :351) and the fp32 -> bf16 operand split that would follow it.
 * x      layout LOCO_NCHW: [R, C, HW] fp32; LOCO_NHWC: [R, HW, C] fp32 or bf16 (dtype), C % 4 == 0.
 * out    [R, ldo] fp32 means (fp32 accumulation, fixed order).
 * hi/lo  optional bf16 operand of the projection GEMM [R, ldh] (hi = rn(mean), lo = rn(mean - hi); lo may be NULL); NULL to skip.
 * loco_spatial_mean_bwd writes dx = dy / HW broadcast over the positions, in the layout / dtype of x. */
int loco_spatial_mean(const void *x, int64_t R, int C, int HW, int layout, int dtype, float *out, int64_t ldo, uint16_t *hi,
                      uint16_t *lo, int64_t ldh, void *stream);
int loco_spatial_mean_bwd(const float *dy, int64_t lddy, int64_t R, int C, int HW, int layout, int dtype, void *dx, void *stream);


/* Inference tail of the box predictor: Detectron2 `fast_rcnn_inference` as reached from `EmbeddingFastRCNNOutputLayers.inference`
 * (reference ovr/modeling/roi_heads/roi_emb_heads.py:280 and :357 -> FastRCNNOutputLayers.inference): Box2BoxTransform.apply_deltas +
 * clip to the image, rows with a non-finite box or score dropped, `score > score_thresh` over the K foreground columns, per-class NMS
 * (torchvision IoU expression, class-agnostic boxes: box_emb_head.py:137), the `topk` best per image.  All images in 4 launches.
 * probs        [R, ld_probs] fp32 probabilities, K + 1 columns used (column K = background, never a candidate).
 * deltas       [R, ld_deltas] fp32 (first 4 columns), proposals [R, 4] fp32 (x1, y1, x2, y2).
 * img_offsets  device int32 [n_img + 1]: rows of image i are [img_offsets[i], img_offsets[i+1]); img_hw device fp32 [n_img, 2] = (h, w).
 * max_rows_per_image  largest image row count (<= 2048).   reg_weights4_host: HOST pointer to (wx, wy, ww, wh).
 * outputs      out_boxes [n_img, topk, 4], out_scores [n_img, topk], out_classes / out_rows int64 [n_img, topk] (out_rows = index of
 *              the RoI among the VALID rows of its image, as the reference returns it), out_count int32 [n_img]; entries beyond
 *              out_count[i] are not written.  Order: score descending, ties by (RoI, class) ascending.
 * workspace    loco_box_inference_workspace_bytes() bytes, 16-byte aligned. */
int64_t loco_box_inference_workspace_bytes(int R, int K, int n_img, int max_rows_per_image, int topk);
int loco_box_inference(const float *probs, int64_t ld_probs, const float *deltas, int64_t ld_deltas, const float *proposals,
                       const int32_t *img_offsets, const float *img_hw, int n_img, int max_rows_per_image, int R, int K,
                       const float *reg_weights4_host, float scale_clamp, float score_thresh, float nms_thresh, int topk,
                       float *out_boxes, float *out_scores, int64_t *out_classes, int64_t *out_rows, int32_t *out_count,
                       void *workspace, void *stream);


/* Multi-token class scoring: the reduction of ``GroundingModule.forward`` (reference ovr/modeling/roi_heads/box_emb_grounding_head.py:
 * 130-221, reached from EmbeddingGroundingFastRCNNOutputLayers.forward_cls_prediction :419-427) after the token GEMM
 * (token_score = Linear(D -> total tokens), :93).  Class k owns the token columns [seg_off[k], seg_off[k+1]) of raw [R, ld_raw]:
 *   s = raw * inv_temp;  a = softmax over the class's tokens (alignment 0) or one-hot of the first maximum (alignment 1);
 *   scores[r, k] = sum_t a_t * s_t     (= -global_dist of the reference; a class without tokens scores 0).
 * att (optional, [R, ld_raw]): the attention weight of every token column (the reference's tok_attention, unpadded).
 * loco_token_pool_bwd: draw[r, t] = dscores[r, k(t)] * d scores / d raw  (softmax Jacobian through the attention; hardmax: the
 * direct term only, as autograd gives for argmax + one_hot). */
int loco_token_pool_fwd(const float *raw, int64_t ld_raw, const int32_t *seg_off, int R, int K1, float inv_temp, int alignment,
                        float *scores, int64_t ld_scores, float *att, void *stream);
int loco_token_pool_bwd(const float *raw, int64_t ld_raw, const int32_t *seg_off, int R, int K1, float inv_temp, int alignment,
                        const float *dscores, int64_t ld_dscores, float *draw, int64_t ld_draw, void *stream);

/* Debug/parity entry: for every roi r, bin (ph,pw) and sample (iy,ix) with iy,ix < max_grid writes
 *   grid_hw [R,2] int32            (gh, gw) — the adaptive sample counts
 *   yx      [R,PH,PW,max_grid,max_grid,2] fp32   unclamped sample coordinate (y, x)
 *   idx     [R,PH,PW,max_grid,max_grid,4] int32  (y_low, x_low, y_high, x_high), -1 = skipped sample
 * Entries with iy >= gh or ix >= gw are left untouched. */
int loco_roi_align_grid_dump(const float *rois, int R, int H, int W, int PH, int PW,
                             float spatial_scale, int sampling_ratio, int aligned, int max_grid,
                             int32_t *grid_hw, float *yx, int32_t *idx, void *stream);

/* ---- operand preparation ------------------------------------------------------------------------
 * fp32 [rows, cols] (row stride src_ld) -> bf16 hi (and lo when lo != NULL), row stride dst_ld.
 * Columns cols..dst_ld-1 of the destination are zero-filled so padded matrices are TMA-safe.
 * transpose != 0 writes dst[c, r] instead (dst is [cols, rows_padded]; dst_ld >= rows).
 * Replaces nothing in the reference (which is fp32-only); it is the fp32 -> tensor-core boundary. */
int loco_split_bf16(const float *src, int64_t rows, int64_t cols, int64_t src_ld, uint16_t *hi,
                    uint16_t *lo, int64_t dst_ld, int transpose, void *stream);


/* loco_lsm_prep with up to three split jobs in the same launch (replaces, for the fp32-accurate LSM step, the separate conversions of
 * the region features and of the v2l_projection weight — reference grounding_head.py:94-111: the masks, `.to(float)` casts and the
 * operands of the projection): job j converts src[j] [rows[j], cols[j]] fp32 (row stride src_ld[j]) into hi[j] / lo[j] bf16
 * [rows[j], dst_ld[j]] (lo[j] may be NULL).  All pointer arrays are HOST arrays of nsplit entries; put the largest job first. */
int loco_lsm_prep_multi(int nsplit, const float *const *src, const int64_t *rows, const int64_t *cols, const int64_t *src_ld,
                        uint16_t *const *hi, uint16_t *const *lo, const int64_t *dst_ld, const int64_t *attention_mask,
                        const int64_t *special_tokens_mask, int64_t n_cap, const void *region_mask, int region_kind, int64_t n_reg,
                        float *cap_mask, float *reg_mask, void *stream);

/* bf16 [rows, cols] (src_ld) -> bf16 [cols, rows] (dst_ld >= rows; pad columns zero-filled): operand
 * re-layout for the backward GEMMs (dX = dY . W needs W^T K-major, dW = dY^T . X needs both transposed). */
int loco_transpose_bf16(const uint16_t *src, int64_t rows, int64_t cols, int64_t src_ld, uint16_t *dst,
                        int64_t dst_ld, void *stream);

/* ---- projection GEMM  out[M,N] = A[M,K] · W[N,K]^T + bias ------------------------------------------
 * Replaces: box_emb_head.py:196,206 (bbox_pred(x), emb_pred(x)) and grounding_head.py:111
 *           (v2l_projection(region_features)) — cuBLAS SGEMM + bias.
 * A_hi/A_lo [M,K] bf16 (lda), W_hi/W_lo [N,K] bf16 (ldw), bias [N] fp32 or NULL.
 * Outputs (each may be NULL): out_f32 [M,N] (ld_f32); out_hi/out_lo [M, n_bf16] bf16 (ld_bf16) holding
 * the first n_bf16 output columns split as hi/lo for a following tensor-core GEMM.
 * tcgen05.mma kind::f16 (bf16 in, fp32 accumulate in TMEM), operands staged by TMA (128B swizzle). */
int loco_linear_fwd(const uint16_t *A_hi, const uint16_t *A_lo, int64_t lda, const uint16_t *W_hi,
                    const uint16_t *W_lo, int64_t ldw, const float *bias, int M, int N, int K,
                    float *out_f32, int64_t ld_f32, uint16_t *out_hi, uint16_t *out_lo,
                    int n_bf16, int64_t ld_bf16, void *stream);

/* Same contraction with the fp32 operands read IN PLACE and multiplied as TF32 (tcgen05 kind::tf32, 10-bit mantissa,
 * fp32 accumulate): no operand-split pass, half the tensor rate.  Used by the reduced-precision ("bf16") mode for
 * the projections whose inputs arrive as fp32 activations.  A [M,K] (lda), W [N,K] (ldw): 16-byte aligned bases,
 * lda % 4 == 0, ldw % 4 == 0. */
int loco_linear_tf32_fwd(const float *A, int64_t lda, const float *W, int64_t ldw, const float *bias, int M,
                         int N, int K, float *out_f32, int64_t ld_f32, uint16_t *out_hi, uint16_t *out_lo,
                         int n_bf16, int64_t ld_bf16, void *stream);

/* ---- RoI x class scoring with fused softmax epilogue -----------------------------------------------
 * Replaces: box_emb_head.py:211 cls_score(e) (cuBLAS) + Detectron2 FastRCNNOutputLayers.predict_probs
 *           F.softmax / .losses log_softmax (ATen) reached from roi_emb_heads.py:266,280,347,357.
 * E_hi/E_lo [R,D] bf16 (lde), C_hi/C_lo [K1,D] bf16 class-embedding matrix (ldc; last row = background),
 * cls_bias [K1] fp32 or NULL.
 * Outputs: logits [R,K1] fp32 (ld_logits) required; probs [R,K1] fp32 or NULL (same ld);
 *          lse [R] fp32 (log-sum-exp over the K1 columns) or NULL;
 *          argmax_fg [R] int64 = argmax over the first K1-1 (foreground) columns, or NULL.
 * Class lists wider than 256 are scored in 256-column chunks dealt to several CTAs per 128-row tile (so that a few
 * thousand RoIs still fill every SM); their partial softmax statistics go through `workspace`
 * (loco_box_score_workspace_bytes(R, K1) bytes, 16-byte aligned; may be NULL when K1 <= 256) and are combined by a second,
 * bandwidth-bound kernel that also writes the probabilities (probs requires lse). */
int64_t loco_box_score_workspace_bytes(int R, int K1);
int loco_box_score_fwd(const uint16_t *E_hi, const uint16_t *E_lo, int64_t lde, const uint16_t *C_hi,
                       const uint16_t *C_lo, int64_t ldc, const float *cls_bias, int R, int K1, int D,
                       float *logits, float *probs, int64_t ld_logits, float *lse,
                       int64_t *argmax_fg, void *workspace, void *stream);

/* Softmax statistics of an EXISTING score matrix (no GEMM in front): the outputs of the loco_box_score_fwd epilogue for
 * logits that did not come out of it.
 * Replaces: Detectron2 FastRCNNOutputLayers.predict_probs F.softmax / .losses F.cross_entropy's log_softmax (ATen) on the
 *           score matrices of box_emb_head.py:204-212 when NORMALIZE_EMB_PRED / STANDARDIZE_EMB_PRED put a row-wise
 *           normalisation between the two GEMMs, or when the caller passes its own scores to losses / inference.
 * logits [R,K1] fp32 (ld_logits).  Outputs (each may be NULL, at least one required): lse [R]; argmax_fg [R] int64 over the
 * first K1-1 columns (first maximum); probs [R,K1] (ld_probs). */
int loco_box_softmax(const float *logits, int64_t ld_logits, int R, int K1, float *lse, int64_t *argmax_fg,
                     float *probs, int64_t ld_probs, void *stream);

/* Row-wise normalisation of the projected embeddings, forward and backward.
 * Replaces: logged_module.py:55-72 normalize_vec (F.normalize p=2) / standardize_vec ((x - mean) / (std + 1e-12), unbiased
 *           std) called at box_emb_head.py:207-210 between emb_pred and cls_score, and at :228-234 on the class matrix.
 * x [rows, cols] fp32 (ldx); mode 0 = L2 normalise, 1 = standardise.  dy == NULL: out = normalised rows.
 * dy [rows, cols] (lddy) given: out = gradient with respect to x (x is the forward input). */
int loco_row_normalize(const float *x, int64_t ldx, int rows, int cols, int mode, const float *dy, int64_t lddy,
                       float *out, int64_t ldo, void *stream);

/* Cross-entropy over the scored logits (Detectron2 FastRCNNOutputLayers.losses: F.cross_entropy mean).
 * labels [R] int64 in [0,K1).  loss_sum: 1 fp32, accumulated (caller zero-fills) with sum_r(lse_r -
 * logit[r,label_r]) * scale.  dlogits (may be NULL): [R,K1] (softmax - onehot) * grad_scale written
 * as fp32 (dlogits_f32, ld_logits) and/or bf16 hi (dlogits_bf16, ld_bf16; zero padded). */
int loco_box_ce_fwd_bwd(const float *logits, int64_t ld_logits, const float *lse, const int64_t *labels,
                        int R, int K1, float scale, float *loss_sum, float grad_scale,
                        float *dlogits_f32, uint16_t *dlogits_bf16, int64_t ld_bf16, void *stream);

/* Class-agnostic box-regression loss and its gradient in one launch.
 * Replaces: Detectron2 FastRCNNOutputLayers.box_reg_loss reached from roi_emb_heads.py:266,347 (foreground mask, nonzero — a host
 *           synchronisation —, Box2BoxTransform.get_deltas, fvcore smooth_l1_loss(sum) / max(R, 1): ~15 ATen launches forward and as
 *           many in autograd's backward).
 * deltas [R,4] fp32 (ld_deltas): predicted (dx, dy, dw, dh); proposal_boxes, gt_boxes [R,4] fp32 contiguous (x1, y1, x2, y2);
 * labels [R] int64 (foreground: 0 <= label < K); reg_weights4_host: the 4 BBOX_REG_WEIGHTS (HOST pointer);
 * loss[0] = scale * sum_fg smooth_l1_beta(deltas - target) (beta < 1e-5: plain L1), overwritten; ddeltas [R,4] (ld_ddeltas) or NULL:
 * d loss / d deltas (zero rows for background).  workspace: loco_box_reg_loss_workspace_bytes(R) bytes, ZERO-INITIALISED once (left
 * zeroed by every launch); partial sums are added in block order: deterministic. */
int64_t loco_box_reg_loss_workspace_bytes(int R);
int loco_box_reg_loss(const float *deltas, int64_t ld_deltas, const float *proposal_boxes, const float *gt_boxes,
                      const int64_t *labels, int R, int K, const float *reg_weights4_host, float smooth_l1_beta, float scale,
                      float *loss, float *ddeltas, int64_t ld_ddeltas, void *workspace, void *stream);

/* Weight / bias gradient of a linear layer with very few outputs: dw[j, v] = sum_r dy[r, j] * x[r, v], db[j] = sum_r dy[r, j], J <= 8.
 * Replaces: autograd's addmm backward for bbox_pred (box_emb_head.py:196; [4 x R] . [R x 2048] through cuBLAS) — here x is streamed
 *           once from HBM with no transposed copy.  dy [R,J] fp32 (ld_dy), x [R,V] fp32 (ld_x % 4 == 0, 16-byte aligned), dw [J,V]
 * contiguous, db [J] or NULL.  workspace: loco_skinny_grad_workspace_bytes(R, J, V) bytes.  Deterministic (fixed chunk order). */
int64_t loco_skinny_grad_workspace_bytes(int R, int J, int V);
int loco_skinny_grad(const float *dy, int64_t ld_dy, const float *x, int64_t ld_x, int R, int J, int V, float *dw, float *db,
                     void *workspace, void *stream);

/* ---- LSM masks -------------------------------------------------------------------------------------
 * Replaces: grounding_head.py:94-96,101,105-106 (attention_mask * (1 - special_tokens_mask), two .to(float32)).
 * attention_mask / special_tokens_mask: n_cap int64 values; region_mask: n_reg values of kind
 * 0 = uint8, 1 = fp32, 2 = int64.  Writes cap_mask / reg_mask as fp32 0/1. */
int loco_lsm_masks(const int64_t *attention_mask, const int64_t *special_tokens_mask, int64_t n_cap,
                   const void *region_mask, int region_kind, int64_t n_reg, float *cap_mask,
                   float *reg_mask, void *stream);

/* ---- LSM input preparation (masks + caption operand in one launch) ---------------------------------------
 * Replaces: grounding_head.py:94-96,101,105-106 (as loco_lsm_masks) and the fp32 -> bf16 (hi / lo) conversion of the caption
 * word embeddings input_caption[TEXT_INPUT] [rows = B*T, cols = D] that the pair GEMM consumes (grounding_head.py:116-147
 * feeds them to torch.bmm as fp32).  cap_lo may be NULL (reduced-precision mode).  dst_ld % 8 == 0, pad columns zeroed. */
int loco_lsm_prep(const float *cap, int64_t rows, int64_t cols, int64_t cap_ld, uint16_t *cap_hi, uint16_t *cap_lo,
                  int64_t dst_ld, const int64_t *attention_mask, const int64_t *special_tokens_mask, int64_t n_cap,
                  const void *region_mask, int region_kind, int64_t n_reg, float *cap_mask, float *reg_mask,
                  void *stream);

/* ---- LSM pair scoring ------------------------------------------------------------------------------
 * Replaces: grounding_head.py:116-256 — B^2 .repeat() replication, torch.bmm, torch.where mask fill,
 *           two F.softmax passes, masked weighted sums (all ATen/cuBLAS launches over [B^2,T,Rg]).
 * cap_hi/cap_lo [Bc*T, D] bf16 (ldcap): caption word embeddings; cap_mask [Bc,T] fp32 (0/1) =
 *           attention_mask * (1 - special_tokens_mask);
 * emb_hi/emb_lo [Bi*Rg, D] bf16 (ldemb): projected region embeddings; reg_mask [Bi,Rg] fp32 (0/1).
 * Writes d_w2r, d_r2w [Bc, Bi] fp32 (row stride ld_out) — rows = captions, cols = images:
 *   d_w2r[c,i] = sum_t m_c[t] sum_r softmax_r(S~)[t,r] * (-S[t,r]) / max(n_words_c, 1)
 *   d_r2w[c,i] = sum_r m_i[r] sum_t softmax_t(S~)[t,r] * (-S[t,r]) / max(n_regions_i, 1)
 * with S = <cap, emb> * inv_temperature and S~ = S where both masks are set, else a finite
 * very-negative fill (an all-masked row/column therefore yields a uniform softmax, as the reference).
 * Either output may be NULL (ALIGN_WORDS_TO_REGIONS / ALIGN_REGIONS_TO_WORDS off).
 * The [B^2,T,Rg] similarity tensor never leaves the SM (TMEM -> registers -> two scalars per pair).
 * Limits: T <= 128, Rg <= 256 (else LOCO_E_UNSUPPORTED).
 * workspace: loco_lsm_pair_workspace_bytes(Bc,T,Bi,Rg) bytes, 16-byte aligned.  RESERVED: the query returns 16 today (every reduction
 *   of the kernel happens on-chip); callers pass a valid pointer so that a later revision may use scratch without an ABI change. */
int64_t loco_lsm_pair_workspace_bytes(int Bc, int T, int Bi, int Rg);
int loco_lsm_pair_fwd(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap,
                      const float *cap_mask, const uint16_t *emb_hi, const uint16_t *emb_lo,
                      int64_t ldemb, const float *reg_mask, int Bc, int T, int Bi, int Rg, int D,
                      float inv_temperature, int alignment, float *d_w2r, float *d_r2w,
                      int64_t ld_out, void *workspace, void *stream);

/* Backward of loco_lsm_pair_fwd with respect to the raw similarities <cap, emb>.
 * Replaces: the autograd graph PyTorch records for grounding_head.py:147-236 (bmm / where / softmax x2 /
 *           mul / sum backward kernels over [B^2,T,Rg]).
 * g_w2r / g_r2w [Bc,Bi] fp32 (ld_g; either may be NULL): upstream gradients of the two distance matrices.
 * The kernel recomputes the similarity tile (same GEMM as forward), evaluates
 *   dS[c,t,i,r] = g_w2r[c,i] * d(d_w2r)/dS + g_r2w[c,i] * d(d_r2w)/dS        (softmax Jacobians included; for
 *   hardmax the one-hot attention is piece-wise constant and only the direct term remains, as in autograd)
 * on-chip and writes it as bf16 hi (+lo when dst_lo != NULL) in the layouts the gradient GEMMs consume:
 *   dst  = dS^T [Bi*Rg, Bc*T] (ld_dst >= Bc*T, % 8)  ->  dEmb [Bi*Rg, D] = dS^T . cap   (loco_linear_fwd)
 *   ds   = dS   [Bc*T, Bi*Rg] (ld_ds, optional)      ->  dCap [Bc*T, D]  = dS . emb
 * Every element of the logical matrices is written (no zero-fill needed). */
int loco_lsm_pair_bwd(const uint16_t *cap_hi, const uint16_t *cap_lo, int64_t ldcap,
                      const float *cap_mask, const uint16_t *emb_hi, const uint16_t *emb_lo,
                      int64_t ldemb, const float *reg_mask, int Bc, int T, int Bi, int Rg, int D,
                      float inv_temperature, int alignment, const float *g_w2r, const float *g_r2w,
                      int64_t ld_g, uint16_t *dst_hi, uint16_t *dst_lo, int64_t ld_dst,
                      uint16_t *ds_hi, uint16_t *ds_lo, int64_t ld_ds, void *stream);

/* ---- peer exchange (multi-GPU step of the sharded LSM head) ------------------------------------------------------------------------
 * New versus the reference, whose heads contain no collective: the image-sharded pair matrix of BASELINE configs[3]
 * needs every rank's caption operands on every rank, and every rank's [B, B_loc] distance block in every rank's [B, B] matrix.
 * Replaces: the NCCL all-gathers a torch.distributed implementation of that exchange would launch (one collective + stream hand-offs
 *           per tensor) AND the barrier after them: the producing rank stores its slices straight into every rank's symmetric buffer
 *           over NVLink and the same launch ends with a cross-GPU flag barrier.
 * nseg <= 4 segments (HOST arrays of length nseg): src[k] [rows[k]][row_bytes[k]] (pitch src_pitch[k], local device memory) is written
 * to peer_p + dst_offset[k] (pitch dst_pitch[k]) for each of the n_peers base pointers in the DEVICE array peers_dev (CUDA
 * symmetric-memory allocations mapped into this process, e.g. torch.distributed._symmetric_memory buffer_ptrs_dev; this rank's own
 * buffer included).  All byte counts / offsets / pitches are multiples of 4 (16-byte stores when they all are multiples of 16).
 * mode: bit 0 = store the segments (nseg > 0), bit 1 = SIGNAL: after the stores, set flag word [channel][rank] of every peer, bit 2 =
 * WAIT: spin until every peer has set this rank's flag words [channel][peer], then clear them — when a launch with the WAIT bit
 * completes, every peer's slices of this exchange have landed here.  (flag_peers_dev: DEVICE array of n_peers pointers to symmetric
 * uint32[16][n_peers] buffers, zero-initialised once; ticket_dev: one zero-initialised uint32 of local device memory.)  SIGNAL and WAIT
 * may be split over two launches (store + signal, independent work, then a WAIT-only launch with nseg = 0) so that the wait for the
 * slowest peer overlaps that work.  Consecutive exchanges must alternate channels.  A segment whose source IS this rank's own
 * destination (slice produced in place) is not copied to itself. */
int loco_peer_exchange(int nseg, const void *const *src, const int64_t *src_pitch, const int *rows, const int64_t *row_bytes,
                       const int64_t *dst_pitch, const int64_t *dst_offset, const void *const *peers_dev, int n_peers,
                       const void *const *flag_peers_dev, int rank, int channel, int mode, void *ticket_dev, void *stream);

/* ---- pair-matrix losses ---------------------------------------------------------------------------
 * Replaces: grounding_head.py:240-251 (empty-pair guard), :272-290 (4 CE losses), :354-379 (accuracies).
 * pw: nmat pair matrices [Bc, Bi] fp32 (row stride ld, matrix stride mat_stride elements; nmat = 2
 * evaluates the w2r and r2w matrices in one launch), each modified IN PLACE by the empty-pair guard:
 *   pw[c,i] = max(pw) + 100 where n_words[c] == 0 and n_regions[i] == 0.
 * cap_mask [Bc,T], reg_mask [Bi,Rg] as above.  The matching caption of image column i is row
 * i + diag_offset (diag_offset = first global image index of this shard when columns are sharded).
 * out4 [nmat,4] fp32 per matrix:
 *             { CE choose-caption (mean over the Bi columns of -log_softmax(-pw, dim=0)[i+off, i]),
 *               CE choose-image   (mean over rows c in [diag_offset, diag_offset+Bi) of
 *                                  -log_softmax(-pw, dim=1)[c, c-off]; only meaningful when Bi == Bc),
 *               accuracy choose-caption, accuracy choose-image }.
 * dpw_caption / dpw_image (each may be NULL): dense [nmat, Bc, Bi] fp32 gradients of out4[.,0] / out4[.,1]
 * with respect to pw (zero at guard-filled entries, which are constants in the reference: .detach()).
 * workspace: loco_pair_ce_workspace_bytes() bytes, needed when max(Bc, Bi) > 32 (several CTAs per matrix, partial sums
 * combined in CTA order by the last one).  It must be zero-initialised ONCE; every launch leaves it zeroed again. */
int64_t loco_pair_ce_workspace_bytes(int nmat, int Bc, int Bi);
int loco_pair_ce(float *pw, int nmat, int64_t mat_stride, int64_t ld, int Bc, int Bi, int diag_offset,
                 const float *cap_mask, int T, const float *reg_mask, int Rg, float *out4,
                 float *dpw_caption, float *dpw_image, void *workspace, void *stream);

/* ---- distillation losses on the pair matrices ------------------------------------------------------------------------------------
 * Replaces: distill_mmss_gcnn.py:226-289 MultiDistillLoss.forward (6 softmax / log_softmax over [B,B], 4 nn.KLDivLoss, transposes) and
 *           :381-433 MultiDistillLossL2.forward (4 nn.MSELoss), called three times per training step at
 *           distill_prop_mmss_gcnn.py:424-442, plus autograd's backward of each: here one launch for the loss and one for all gradients.
 * trans (teacher), w2r, r2w: [B,B] fp32 pair matrices (rows = captions, columns = images; row strides ld_*).
 * kind 0: KD with the teacher as target (DISTILLATION_TEACHER_TRANSFORMER = True); 1: KD with the students as targets (False, the shipped
 * coco_lsm.yaml:63); 2: MSE.  loss[0] = weighted loss (overwritten).  g_trans / g_w2r / g_r2w: dense [B,B] gradients of the loss, each may
 * be NULL (a detached side, DISTILLATION_DETACH_TEACHER, or no gradient wanted).  workspace: loco_pair_distill_workspace_bytes(B)
 * bytes, ZERO-INITIALISED once (left zeroed by every launch).  Deterministic (fixed summation order). */
int64_t loco_pair_distill_workspace_bytes(int B);
int loco_pair_distill(const float *trans, int64_t ld_trans, const float *w2r, int64_t ld_w2r, const float *r2w, int64_t ld_r2w, int B,
                      float temperature, int kind, float loss_weight, float *loss, float *g_trans, float *g_w2r, float *g_r2w,
                      void *workspace, void *stream);

/* ---- device-side statistics of a logged tensor -----------------------------------------------------------------------------------
 * Replaces: logged_module.py:8-18 stats() — tensor.cpu() plus four scalar reductions with a host synchronisation each — called by
 *           LoggedModule.log for every logged tensor (seven per GroundingHead.forward, grounding_head.py:97-114,158,253-255).
 * x: n fp32 values (contiguous).  out4 (device) = {min, max, mean, std (unbiased, as torch.Tensor.std)}.  workspace:
 * loco_tensor_stats_workspace_bytes() bytes, 8-byte aligned, ZERO-INITIALISED once (left zeroed).  Nothing is copied to the host. */
int64_t loco_tensor_stats_workspace_bytes(void);
int loco_tensor_stats(const float *x, int64_t n, float *out4, void *workspace, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* LOCOV_B200_H */
