/*
 * scda_b200.h — C ABI of libscda_b200.so, the sm_100a operator library behind
 * the SCDA Faster R-CNN hot path.
 *
 * Boundary rules (all entry points):
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless
 *     the parameter name ends in _host;
 *   - the caller allocates every output (and the workspace, sized by the
 *     matching *_workspace_bytes query); nothing is allocated or freed inside;
 *   - work is enqueued on the cudaStream_t passed in and the call returns
 *     without synchronising; re-entrant, and no global state apart from the two
 *     process-wide tuning / test hooks (scda_conv3x3_set_plan,
 *     scda_conv3x3_wgrad_set_form), which select between equivalent kernels;
 *   - return value: 1 = launched, 0 = rejected arguments, negative =
 *     -(cudaError_t) of the failed launch.  Never exit()s (the reference
 *     launchers do: extensions/_roi_pooling/src/roi_pooling_kernel.cu:117-122).
 *
 * Section A keeps the names and parameter lists of the reference's own
 * launcher layer (sic "Laucher"), one level below its TH/THC cffi glue which
 * no longer exists in torch >= 1.0, so the unmodified reference objects
 * (oracle/_ref/libscda_ref.so) and this library are interchangeable under one
 * binding.  Section B adds the fused / on-device forms the B200 path uses.
 * All file:line citations are relative to the reference tree.
 */
#ifndef SCDA_B200_H_
#define SCDA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __DRIVER_TYPES_H__
typedef struct CUstream_st *cudaStream_t;
#endif

/* ===================================================================== */
/* A. Reference launcher layer, same symbols                               */
/* ===================================================================== */

/* replaces extensions/_roi_pooling/src/roi_pooling_kernel.h:8-12
 * (behind roi_pooling_forward_cuda, src/roi_pooling_cuda.c:7-38).
 * bottom_data NCHW fp32; bottom_rois [num_rois,5] = (batch, x1, y1, x2, y2)
 * in image coordinates; top_data [num_rois,C,PH,PW]; argmax_data same shape,
 * flat offset into bottom_data of the max (-1 for an empty bin), may be NULL. */
int ROIPoolForwardLaucher(const float *bottom_data, const float spatial_scale, const int num_rois,
                          const int height, const int width, const int channels,
                          const int pooled_height, const int pooled_width,
                          const float *bottom_rois, float *top_data, int *argmax_data,
                          cudaStream_t stream);

/* replaces roi_pooling_kernel.h:15-18 (behind roi_pooling_backward_cuda,
 * src/roi_pooling_cuda.c:41-88).  bottom_diff [batch,C,H,W] is overwritten. */
int ROIPoolBackwardLaucher(const float *top_diff, const float spatial_scale, const int batch_size,
                           const int num_rois, const int height, const int width,
                           const int channels, const int pooled_height, const int pooled_width,
                           const float *bottom_rois, float *bottom_diff, const int *argmax_data,
                           cudaStream_t stream);

/* replaces extensions/_roi_align/src/roi_align_kernel.h:13-17
 * (behind roi_align_forward_cuda, src/roi_align_cuda.c:7-38). */
int ROIAlignForwardLaucher(const float *bottom_data, const float spatial_scale, const int num_rois,
                           const int height, const int width, const int channels,
                           const int aligned_height, const int aligned_width,
                           const float *bottom_rois, float *top_data, cudaStream_t stream);

/* replaces roi_align_kernel.h:24-27 (behind roi_align_backward_cuda,
 * src/roi_align_cuda.c:40-76).  Accumulates INTO bottom_diff, which the
 * caller zeroes (functions/roi_align.py:40-41), like the reference. */
int ROIAlignBackwardLaucher(const float *top_diff, const float spatial_scale, const int batch_size,
                            const int num_rois, const int height, const int width,
                            const int channels, const int aligned_height, const int aligned_width,
                            const float *bottom_rois, float *bottom_diff, cudaStream_t stream);

/* replaces extensions/_nms/src/cuda/nms_kernel.h:11-12: the N x ceil(N/64)
 * suppression bitmask on the legacy default stream (the reference passes no
 * stream, nms_kernel.cu:79).  Lower-triangle words are written as 0. */
void _nms(int boxes_num, float *boxes_dev, unsigned long long *mask_dev, float nms_overlap_thresh);

/* replaces extensions/_bbox_helper/src/cuda/iou_overlap_kernel.h:8-14
 * (behind gpu_iou_overlaps, src/bbox_helper_cuda.c:16-47). */
int IOUOverlap(const float *bboxes1_data, const float *bboxes2_data, const int size_bbox,
               const int num_bbox1, const int num_bbox2, float *top_data, cudaStream_t stream);

/* replace extensions/_focal_loss/src/cuda/focal_loss_sigmoid_kernel.h:8-18
 * (behind focal_loss_sigmoid_{forward,backward}_cuda,
 * src/focal_loss_cuda.c:10-53).  N = rows * num_classes. */
int SigmoidFocalLossForwardLaucher(const int N, const float *logits, const int *targets,
                                   const float weight_pos, const float gamma, const float alpha,
                                   const int num_classes, float *losses, cudaStream_t stream);
int SigmoidFocalLossBackwardLaucher(const int N, const float *logits, const int *targets,
                                    float *dX_data, const float weight_pos, const float gamma,
                                    const float alpha, const int num_classes, cudaStream_t stream);

/* replace focal_loss_softmax_kernel.h:8-19 (behind
 * focal_loss_softmax_{forward,backward}_cuda, src/focal_loss_cuda.c:55-108). */
int SoftmaxFocalLossForwardLaucher(const int N, const float *logits, const int *targets,
                                   const float weight_pos, const float gamma, const float alpha,
                                   const int num_classes, float *losses, float *priors,
                                   cudaStream_t stream);
int SoftmaxFocalLossBackwardLaucher(const int N, const float *logits, const int *targets,
                                    float *dX_data, const float weight_pos, const float gamma,
                                    const float alpha, const int num_classes, const float *priors,
                                    float *buff, cudaStream_t stream);

/* ===================================================================== */
/* B. B200 path                                                            */
/* ===================================================================== */

int scda_abi_version(void);

/* --- NMS, wholly on device -------------------------------------------- */
/* replaces gpu_nms (extensions/_nms/src/nms_cuda.c:17-67): bitmask kernel +
 * D2H of the mask + sequential host scan become bitmask kernel (upper
 * triangle only) + single-CTA device scan; the mask never leaves HBM/L2.
 * boxes [n,5] = (x1,y1,x2,y2,score) sorted by descending score; keep_out
 * int64[n] receives ascending kept indices, num_out int64[1] their count.
 * max_keep > 0 stops the scan after that many survivors (what the callers'
 * keep[:post_nms_top_n] slicing discards, functions/rpn_proposal.py:65-66);
 * 0 = scan everything.  workspace >= scda_nms_workspace_bytes(n). */
size_t scda_nms_workspace_bytes(int n);
int scda_nms(int n, const float *boxes, float thresh, int max_keep, int64_t *keep_out,
             int64_t *num_out, void *workspace, size_t workspace_bytes, cudaStream_t stream);
/* same, with the live box count read from device memory (int32 n_dev[0], clamped to
 * [0, n_cap]): lets a pipeline whose earlier stages filter boxes on the device
 * (min-size filter, functions/rpn_proposal.py:62-63) run NMS without a host
 * round trip for the count.  Buffers and workspace are sized for n_cap. */
int scda_nms_dyn(int n_cap, const int *n_dev, const float *boxes, float thresh, int max_keep,
                 int64_t *keep_out, int64_t *num_out, void *workspace, size_t workspace_bytes,
                 cudaStream_t stream);
/* the bitmask alone, on a stream (the _nms symbol above is the stream-less form) */
/* `groups` independent NMS problems in one launch (the eight per-class calls per image of
 * compute_predicted_bboxes, functions/predict_bbox.py:36-52): boxes [groups][n_cap][5] each sorted by
 * descending score, n_dev[g] <= n_cap <= 1024 live boxes; keep_out [groups][n_cap] ascending survivor
 * indices, num_out [groups].  Same survivors as scda_nms on each group. */
int scda_nms_groups(int groups, int n_cap, const int *n_dev, const float *boxes, float thresh,
                    int64_t *keep_out, int64_t *num_out, cudaStream_t stream);
int scda_nms_mask(int n, const float *boxes, unsigned long long *mask, float thresh,
                  cudaStream_t stream);

/* --- final detections of one image (functions/predict_bbox.py:13-66) ------------------- */
/* Front half: per foreground class c = 1 .. num_classes-1 the score threshold (:39-42; <= 0: none), the
 * descending-score order (:44-45) and the float64 decode + clip of the class's deltas (:26-31,38) ->
 * dets [num_classes-1][n][5] (x1, y1, x2, y2, score; rows behind n_live[c-1] carry score -1), the input
 * layout of scda_nms_groups.  rois [n][roi_stride] (batch, x1, y1, x2, y2, ...), cls [n][num_classes]
 * probabilities, loc [n][4*num_classes]; stds / means: 4 host doubles each (normalize != 0). n <= 1024. */
int scda_predict_prepare(int n, int num_classes, const float *rois, int roi_stride, const float *cls,
                         const float *loc, int normalize, const double *stds, const double *means,
                         double img_h, double img_w, float score_thresh, float *dets, int *n_live,
                         cudaStream_t stream);
/* Back half (:53-63): the survivors of all classes (keep / n_keep as written by scda_nms_groups) ranked by
 * descending score -> rows [top_n][7] (batch, x1, y1, x2, y2, score, class) and count[0] = rows written. */
int scda_predict_topn(int groups, int n, const float *dets, const int64_t *keep, const int64_t *n_keep,
                      float batch_ix, int top_n, float *rows, int *count, cudaStream_t stream);

/* --- IoU, cython_bbox convention -------------------------------------- */
/* replaces cython_bbox.bbox_overlaps (extensions/_cython_bbox/cython_bbox.pyx:32-73,
 * reached through utils/bbox_helper.py:8-9): boxes [n,4], query [k,4] -> out [n,k];
 * no +1, zero unless both overlaps > 0, no clamp.  Bit-exact with the host code. */
int scda_bbox_overlaps(int n, const float *boxes, int k, const float *query, float *out,
                       cudaStream_t stream);

/* --- focal losses with the reduction fused ---------------------------- */
/* SigmoidFocalLossFunction.forward = kernel + losses.sum() in Python
 * (extensions/_focal_loss/focal_loss.py:31-46): one pass, loss_sum[0] is
 * overwritten with the total.  losses may be NULL (skip the M x K write). */
int scda_sigmoid_focal_loss_sum(const int N, const float *logits, const int *targets,
                                const float weight_pos, const float gamma, const float alpha,
                                const int num_classes, float *losses, float *loss_sum,
                                cudaStream_t stream);
/* SoftmaxFocalLossFunction.forward (focal_loss.py:103-118), same fusion;
 * priors [N] is still written (backward needs it), losses may be NULL. */
int scda_softmax_focal_loss_sum(const int N, const float *logits, const int *targets,
                                const float weight_pos, const float gamma, const float alpha,
                                const int num_classes, float *losses, float *priors,
                                float *loss_sum, cudaStream_t stream);

/* --- RPN proposal decode ------------------------------------------------ */
/* the stretch of compute_rpn_proposals between the top-k and the NMS (functions/rpn_proposal.py:53-64,
 * utils/bbox_helper.py:88-111): rows i < pre take anchor / delta order[i]; decode in the reference's dtypes
 * (float32 exp, float64 products, every operation rounded on its own), clip to [0, img_w - 1] x [0, img_h - 1],
 * drop boxes with w + 1 or h + 1 < min_size, compact the survivors in order into packed[pre][5] =
 * (x1, y1, x2, y2, score) float32 (zero rows behind them) and write their number to count[0].
 * anchors [KA][4] float64, deltas [KA][4] float32, order [pre] int64, top_scores [pre]. */
int scda_rpn_decode_pack(int pre, const double *anchors, const float *deltas, const long long *order,
                         const float *top_scores, double img_h, double img_w, double min_size, float *packed,
                         int *count, cudaStream_t stream);

/* the same rows straight from the scores: the top-`pre` selection and the descending sort of
 * functions/rpn_proposal.py:49-55 (numpy argpartition + argsort on the host in the reference) run on the device
 * too.  scores [KA] float32 (foreground probability per anchor); pre <= 0 or > KA takes every anchor.  Equal
 * scores rank by ascending anchor index (the reference leaves that order to numpy's unstable sort).
 * KA <= 51 200.  workspace: scda_rpn_proposal_rows_workspace_bytes(KA, pre) bytes, 16-byte aligned; two
 * launches, no host synchronisation. */
size_t scda_rpn_proposal_rows_workspace_bytes(int KA, int pre);
int scda_rpn_proposal_rows(int KA, int pre, const float *scores, const double *anchors, const float *deltas,
                           double img_h, double img_w, double min_size, float *packed, int *count,
                           void *workspace, size_t workspace_bytes, cudaStream_t stream);

/* dst[c][r] = src[r][c], bf16 [rows][cols] -> [cols][rows]; lds / ldd: leading dimensions in elements */
int scda_transpose_bf16(int rows, int cols, const void *src, long long lds, void *dst, long long ldd,
                        cudaStream_t stream);

/* --- runtime ------------------------------------------------------------ */
/* identity of the stream capture `stream` is part of, 0 when it is not capturing (cudaStreamGetCaptureInfo) */
unsigned long long scda_stream_capture_id(cudaStream_t stream);
/* buf[slot] = %globaltimer (ns) in stream order: phase boundaries measured from inside a replayed graph */
int scda_timestamp(unsigned long long *buf, int slot, cudaStream_t stream);

/* --- RCNN proposal targets --------------------------------------------- */
/* compute_proposal_targets for one image (functions/proposal_target.py:17-177): boxes float32 [cap][ldb >= 4]
 * = (x1, y1, x2, y2, ...) of which the first n_boxes[0] (device int64) are live; gts float32 [G][5] =
 * (x1, y1, x2, y2, label), rows with x2 <= x1 + 1 or y2 <= y1 + 1 are padding; append_gts adds the valid gts to
 * the candidates (:42-43).  Clip to the image, IoU (the cython_bbox arithmetic), positives iou > pos_thresh,
 * negatives neg_lo <= iou < neg_hi; at most want_pos positives, the rest negatives, up to batch_size rows,
 * padded by resampling.  Draws are key driven (functions/_sampling.py): candidate j (ascending index) owns
 * keys[j], the k candidates with the smallest keys are chosen in key order; keys_pos / keys_neg [>= R],
 * keys_pad [>= batch_size] float64 in [0, 1) on the device, R = cap (+ G).  Outputs (device): rois
 * [batch_size][5] = (batch_ix, box), labels [batch_size] int64, loc_targets / loc_weights
 * [batch_size][4 * num_classes] float32, class-specific, (t - mean) / std when `normalize` (means4 / stds4: HOST
 * pointers to four doubles).  cap + G <= 4096, G <= 256.  One launch, no host synchronisation. */
int scda_proposal_targets(int cap, int ldb, const float *boxes, const long long *n_boxes, int G, const float *gts,
                          int append_gts, float img_h, float img_w, float pos_thresh, float neg_hi, float neg_lo,
                          int want_pos, int batch_size, int num_classes, int normalize, const double *means4,
                          const double *stds4, float batch_ix, const double *keys_pos, const double *keys_neg,
                          const double *keys_pad, float *rois, long long *labels, float *loc_targets,
                          float *loc_weights, cudaStream_t stream);

/* --- RPN anchor targets ------------------------------------------------- */
/* compute_anchor_targets for one image (functions/anchor_target.py:16-116): anchors32 float32 / anchors64
 * float64 [fh*fw*A][4] (the same table, cell-major then anchor), gts float32 [G][5] (zero rows = padding).
 * IoU (cython_bbox arithmetic), label 0 below neg_thresh, 1 above pos_thresh or where an anchor attains a
 * ground truth's maximum (>= 0.1, ties kept, the largest ground-truth index wins the match), -1 otherwise; at
 * most want_pos positives and batch_total labelled anchors survive, the rest go back to -1 — key-driven draws
 * as scda_proposal_targets (keys_pos / keys_neg [>= fh*fw*A] float64: the members with the SMALLEST keys are
 * removed).  Outputs: cls_targets int64 [A][fh][fw], loc_targets / loc_masks float32 [4A][fh][fw] (encode in
 * float64 without +1 widths, utils/bbox_helper.py:60-85), normalizer[0] = max(1, labelled anchors).
 * G <= 256, fh*fw*A <= 100 000.  One launch, no host synchronisation. */
int scda_anchor_targets(int A, int fh, int fw, const float *anchors32, const double *anchors64, int G,
                        const float *gts, float neg_thresh, float pos_thresh, int want_pos, int batch_total,
                        const double *keys_pos, const double *keys_neg, long long *cls_targets, float *loc_targets,
                        float *loc_masks, long long *normalizer, cudaStream_t stream);

/* crops around the cluster centres (tools/faster_rcnn_train_val.py:411-438, 528-557): K windows of R x R pixels,
 * corner = clamp(int(centre) - R/2, 0, size - R) per axis, gathered from image [C, H, W] fp32 into
 * out [K, C, R, R]; centers [K][2] = (x, y) fp32 on the device.  R even, R <= H, W. */
int scda_crop_regions(int K, int C, int H, int W, int R, const float *image, const float *centers, float *out,
                      cudaStream_t stream);

/* --- fused detector / adversarial losses -------------------------------- */
/* smooth_l1_loss_with_sigma(pred * mask, target) of the reference
 * (models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:238-246; RPN :54-55, RCNN :64-66):
 * loss_sum[0] = sum_i f(pred[i] * mask[i] - target[i]),  f(d) = 0.5 sigma^2 d^2 if |d| < 1/sigma^2 else
 * |d| - 0.5/sigma^2.  mask may be NULL (= 1).  One block, fixed summation order (deterministic). */
int scda_smooth_l1_sigma_sum_fwd(long long n, const float *pred, const float *mask, const float *target,
                                 float sigma, float *loss_sum, cudaStream_t stream);
/* grad_pred[i] = grad_loss[0] * f'(d_i) * mask[i] */
int scda_smooth_l1_sigma_sum_bwd(long long n, const float *pred, const float *mask, const float *target,
                                 float sigma, const float *grad_loss, float *grad_pred, cudaStream_t stream);
/* the adversarial terms of the four-phase update (tools/faster_rcnn_train_val.py:577-600, 655-680, 716-732):
 * row_mean[k] = mean_m F.binary_cross_entropy(sigmoid(logits[k, m]), labels[m * label_stride]) for K cluster
 * rows against ONE label row (label_stride 1) or one constant (label_stride 0); logs clamped at -100 as
 * torch does.  One block per row, fixed summation order. */
int scda_bce_sigmoid_rows_fwd(int K, int M, const float *logits, const float *labels, int label_stride,
                              float *row_mean, cudaStream_t stream);
/* grad_logits[k, m] = grad_rows[k] / M * (p - y) / max(p (1 - p), 1e-12) * p (1 - p),  p = sigmoid(logit) */
int scda_bce_sigmoid_rows_bwd(int K, int M, const float *logits, const float *labels, int label_stride,
                              const float *grad_rows, float *grad_logits, cudaStream_t stream);

/* decoder head: nn.ConvTranspose2d(Cin = 32, Cout <= 4, kernel_size = 1) + nn.Tanh, the last two layers of each
 * decoder (models/faster_rcnn/faster_rcnn_adver_expansion_reweight_cluster.py:380-383), one streaming pass each
 * way.  x [P, Cin], y / dy [P, Cout], dx [P, Cin] fp32 with the channel innermost (channels-last); w [Cin][Cout]
 * (the ConvTranspose2d weight [Cin, Cout, 1, 1] as stored); b [Cout] or NULL; dx may be NULL; dw [Cin][Cout],
 * db [Cout] (or NULL) are overwritten.  Reductions in a fixed order (deterministic). */
size_t scda_conv1x1_tanh_workspace_bytes(long long P, int Cin, int Cout);
int scda_conv1x1_tanh_fwd(long long P, int Cin, int Cout, const float *x, const float *w, const float *b,
                          float *y, cudaStream_t stream);
int scda_conv1x1_tanh_bwd(long long P, int Cin, int Cout, const float *x, const float *w, const float *y,
                          const float *dy, float *dx, float *dw, float *db, void *workspace,
                          size_t workspace_bytes, cudaStream_t stream);

/* --- tensor-core GEMM / 3x3 convolution (tcgen05 + TMA) ----------------- */
/* flags for both entry points */
#define SCDA_TC_RELU        1   /* y = max(y, 0)                                        */
#define SCDA_TC_OUT_F32     2   /* output fp32 (default bf16)                            */
#define SCDA_TC_MASK_POS    4   /* y = mask_src > 0 ? y : 0  (ReLU backward, bf16 mask)  */
#define SCDA_TC_ACCUMULATE  8   /* y += old y (fp32 output only)                         */
#define SCDA_TC_MUL_SRC     16  /* y *= mul_src (bf16, e.g. the dropout keep/scale tensor) */
/* replaces the cuBLAS GEMM behind nn.Linear (fc6 / fc7 / fc_rcnn_cls / fc_rcnn_loc,
 * models/faster_rcnn/vgg_adver_expansion_cluster.py:46-60,73-80) and behind 1x1 convs
 * (models/head.py:15-18):  C[M,N] = A[M,K] . B[N,K]^T + bias[N].
 * A, B bf16 row-major with leading dimensions lda, ldb (elements, multiples of 8, 16 B
 * aligned bases); C bf16 or fp32 with leading dimension ldc; bias fp32 or NULL. */
int scda_gemm_bf16_tn(int M, int N, int K, const void *A, long long lda, const void *B, long long ldb,
                      const float *bias, void *C, long long ldc, int flags, const void *mask_src,
                      const void *mul_src, cudaStream_t stream);
/* replaces the cuDNN convolution behind nn.Conv2d(k=3, s=1, p=1) (VGG backbone,
 * vgg_adver_expansion_cluster.py:101-114; RPN conv3x3, models/head.py:13):
 * x NHWC bf16 [NB,H,W,Cin], w bf16 [Cout][3][3][Cin], y NHWC [NB,H,W,Cout] bf16 or fp32.
 * Cin % 64 == 0, Cout % 32 == 0 (32 output channels ride in a 64-wide tile whose upper weight rows are TMA
 * out-of-bounds zeros), W % 8 == 0.  With the weights flipped and transposed by the caller the same
 * entry point computes the data gradient; SCDA_TC_MASK_POS then applies the ReLU gradient of
 * the layer input (mask_src = that input, same shape as y). */
int scda_conv3x3_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *x, const void *w_krsc,
                           const float *bias, void *y, int flags, const void *mask_src,
                           cudaStream_t stream);

/* data gradient of that convolution read straight from the FORWARD weights (no flipped /
 * transposed copy is kept): dy NHWC bf16 [NB,H,W,Cout], w bf16 [Cout][3][3][Cin] ->
 * dx NHWC [NB,H,W,Cin] (bf16, or fp32 with SCDA_TC_OUT_F32).  The weights enter the MMA as an
 * MN-major operand with the tap mirrored.  SCDA_TC_MASK_POS applies the ReLU gradient of the
 * layer input (mask_src = that input).  Cin % 64 == 0, Cout % 32 == 0 (halo form; the per-tap form needs
 * Cout % 64 == 0), W % 8 == 0. */
int scda_conv3x3_dgrad_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *dy,
                                 const void *w_krsc, void *dx, int flags, const void *mask_src,
                                 cudaStream_t stream);

/* First convolution of the backbone straight from the fp32 NCHW image (models/faster_rcnn/
 * vgg_adver_expansion_cluster.py:101-114, features[0] + ReLU): y[n,h,w,co] = relu?(bias[co] +
 * sum x[n,ci,h+kh-1,w+kw-1] * w[co][kh][kw][ci]).  x fp32 [NB,Cin,H,W] (Cin <= 3; rounded to bf16 as it is
 * staged), w bf16 [64][3][3][Cin], y bf16 NHWC [NB,H,W,64], fp32 accumulation.  The im2col exists only in
 * shared memory (csrc/conv_first.cu); Cout must be 64. */
int scda_conv3x3_first_nchw(int NB, int H, int W, int Cin, int Cout, const float *x, const void *w_krsc,
                            const float *bias, void *y, int relu, cudaStream_t stream);

/* Tuning / test hook for the two 3x3 entry points above: halo = 1 stages the input tile once for
 * all nine taps (csrc/conv_halo.cu, the default), 0 = one TMA box per tap (csrc/gemm_tc.cu);
 * block_n (0 auto | 64 | 128) and sub_tiles (0 auto | 1 | 2) force the halo kernel's tile plan.
 * A negative argument leaves that setting unchanged.  Results are identical up to fp32
 * summation order.  Process-wide, not synchronised with concurrent launches. */
int scda_conv3x3_set_plan(int halo, int block_n, int sub_tiles);

/* CTA-pair form of the halo kernel (tcgen05 cta_group::2: two SMs run one 256-pixel x block_n MMA per
 * k-step and each stages half of every weight tile): -1 = the measured per-layer plan (default), 0 = never,
 * 1 = wherever legal (stride 1; forward block_n 64 | 128 | 256, data gradient 128 | 256).  With pairs a
 * block_n of 256 may be forced through scda_conv3x3_set_plan.  Same results up to fp32 summation order. */
int scda_conv3x3_set_pair(int mode);

/* C[M,N] = A[M,K] . B[K,N] + bias: B row-major with N contiguous (the data gradient of
 * nn.Linear, dX = dY . W, reads W[out,in] this way; no transposed weight copy is kept). */
int scda_gemm_bf16_nn(int M, int N, int K, const void *A, long long lda, const void *B, long long ldb,
                      const float *bias, void *C, long long ldc, int flags, const void *mask_src,
                      const void *mul_src, cudaStream_t stream);
/* weight gradient of nn.Linear: dW[Nout,Kin] (fp32, leading dimension lddw) =
 * dY[rows,Nout]^T . X[rows,Kin], both bf16 row-major. */
int scda_linear_wgrad_bf16(int rows, int Nout, int Kin, const void *dY, long long lddy, const void *X,
                           long long ldx, float *dW, long long lddw, int accumulate,
                           cudaStream_t stream);
/* weight gradient of the 3x3 convolution: x NHWC bf16 [NB,H,W,Cin], dy NHWC bf16
 * [NB,H,W,Cout] -> dw_partials fp32 [splits][Cout][3][3][Cin]; the pixel reduction is cut
 * into `splits` equal ranges of 128-pixel tiles (every slab is written; the caller sums
 * them, so the result does not depend on scheduling).  Cin % 64 == 0, Cout % 32 == 0. */
int scda_conv3x3_wgrad_bf16_nhwc(int NB, int H, int W, int Cin, int Cout, const void *x, const void *dy,
                                 float *dw_partials, int splits, cudaStream_t stream);

/* Tuning / test hook: form of scda_conv3x3_wgrad_bf16_nhwc.  0 (default) = one tap per CTA; 1 = one kernel
 * column (three taps) per CTA, the shifted input rows staged once for the three (faster in isolation, but it
 * occupies up to all 512 TMEM columns of the SM).  Same results up to fp32 summation order.  Process-wide. */
int scda_conv3x3_wgrad_set_form(int three_taps);

/* --- bf16 NHWC companions of the tensor-core kernels (HBM bound) --------- */
/* nn.MaxPool2d(2, 2) of the VGG stack (vgg_adver_expansion_cluster.py:105-106) on NHWC bf16 */
int scda_maxpool2x2_nhwc_bf16(int NB, int H, int W, int C, const void *x, void *y, cudaStream_t stream);
/* its backward (gradient to the first maximum of each window, PyTorch's tie rule); with
 * relu_mask != 0 also the backward of the nn.ReLU below it (x = that ReLU's output) */
int scda_maxpool2x2_bwd_nhwc_bf16(int NB, int H, int W, int C, const void *x, const void *dy, void *dx,
                                  int relu_mask, cudaStream_t stream);
/* image NCHW fp32 -> NHWC bf16 with the channel dimension zero-padded to Cpad */
int scda_nchw_f32_to_nhwc_bf16(int NB, int C, int H, int W, int Cpad, const float *x, void *y,
                               cudaStream_t stream);
/* feature map NHWC bf16 -> NCHW fp32 (the layout of the reference's RoI operators) */
int scda_nhwc_bf16_to_nchw_f32(int NB, int C, int H, int W, const void *x, float *y, cudaStream_t stream);
/* dst[n] (+)= sum of n_slabs fp32 slabs (split-K partial weight gradients) */
int scda_reduce_slabs_f32(const float *slabs, long long slab_stride, int n_slabs, float *dst,
                          long long n, int accumulate, cudaStream_t stream);
/* bias gradient: out[N] += column sums of x[M, ld] (bf16) */
int scda_colsum_bf16(long long M, int N, const void *x, long long ld, float *out, cudaStream_t stream);
/* the same for a contiguous fp32 matrix x[M, N] (N % 4 == 0, N / 4 a power of two <= 256): bias
 * gradient of the channels-last convolutions of the reconstruction networks
 * (models/faster_rcnn/common_net.py:59-80, 251-293: nn.Conv2d(bias=True)) */
int scda_colsum_f32(long long M, int N, const float *x, float *out, cudaStream_t stream);

/* --- InstanceNorm2d (+ activation), channels-last fp32 ------------------- */
/* replaces nn.InstanceNorm2d(affine=False) and the ReLU / LeakyReLU behind it in the
 * reconstruction networks (models/faster_rcnn/common_net.py:59-80, 279-293).
 * x, y, dy, dx: [N, HW, C] fp32 (C innermost = channels_last), C % 4 == 0, C <= 1024 and
 * 1024 / C a power of two; mean, rstd: [N, C] (written by fwd, read by bwd).
 * act: 0 none, 1 ReLU, 2 LeakyReLU(slope).  Reductions are two-stage in a fixed order
 * (deterministic).  workspace >= scda_instnorm_workspace_bytes (+ 8*N*C bytes for bwd). */
size_t scda_instnorm_workspace_bytes(int N, int HW, int C);
int scda_instnorm_act_fwd_nhwc_f32(int N, int HW, int C, const float *x, float *y, float *mean, float *rstd,
                                   float eps, int act, float slope, void *workspace,
                                   size_t workspace_bytes, cudaStream_t stream);
int scda_instnorm_act_bwd_nhwc_f32(int N, int HW, int C, const float *x, const float *dy, const float *mean,
                                   const float *rstd, float *dx, int act, float slope, void *workspace,
                                   size_t workspace_bytes, cudaStream_t stream);

/* bilinear x2 up-sampling, align_corners=True (common_net.py:160-169 `Interpolate`), on
 * channels-last fp32: x [N,H,W,C] -> y [N,2H,2W,C]; the backward is its transpose written as
 * a gather (deterministic).  C % 4 == 0. */
int scda_upsample_bilinear2x_nhwc_f32(int N, int H, int W, int C, const float *x, float *y,
                                      cudaStream_t stream);
int scda_upsample_bilinear2x_bwd_nhwc_f32(int N, int H, int W, int C, const float *dy, float *dx,
                                          cudaStream_t stream);

/* the same kernels with the output (fwd: y; bwd: dx) and the incoming gradient (bwd: dy) in fp32
 * (dtype code 0) or bf16 (1): the bf16 forms feed / are fed by the tensor-core convolutions of
 * the decoder without a separate cast pass.  x, mean, rstd stay fp32. */
int scda_instnorm_act_fwd_nhwc(int N, int HW, int C, const float *x, void *y, int y_dtype, float *mean,
                               float *rstd, float eps, int act, float slope, void *workspace,
                               size_t workspace_bytes, cudaStream_t stream);
int scda_instnorm_act_bwd_nhwc(int N, int HW, int C, const float *x, const void *dy, int dy_dtype,
                               const float *mean, const float *rstd, void *dx, int dx_dtype, int act,
                               float slope, void *workspace, size_t workspace_bytes, cudaStream_t stream);
/* bilinear x2 up-sampling (align_corners) of a channels-last fp32 tensor into fp32 (0) or bf16 (1) */
int scda_upsample_bilinear2x_nhwc(int N, int H, int W, int C, const float *x, void *y, int y_dtype,
                                  cudaStream_t stream);
/* its transpose: dy [N, 2H, 2W, C] fp32 (0) or bf16 (1) -> dx [N, H, W, C] fp32 */
int scda_upsample_bilinear2x_bwd_nhwc(int N, int H, int W, int C, const void *dy, int dy_dtype, float *dx,
                                      cudaStream_t stream);

/* --- region grouping (k-means of RoI centres) --------------------------- */
/* replaces, inside compute_cluster_targets (functions/mask.py:193-237), the host call
 * sklearn.cluster.KMeans(n_clusters=k, random_state=0).fit(centres) on the float32 RoI
 * centres (:205-211) and the per-cluster member selection (:213-224).
 * rois [n, roi_stride] fp32 rows (b, x1, y1, x2, y2, ...).  first_center_id and
 * uniforms[(k-1) * n_local_trials] are the data-independent draws of numpy's
 * RandomState(0) that k-means++ consumes (the caller generates them once per n, k).
 * Outputs: labels [n] int32, centers [k, 2] fp32 (cx, cy), counts [k] int32 and, when
 * threshold > 0, index [k * threshold] int64 = for each cluster its first `threshold`
 * members in ascending RoI order, or — for a smaller cluster — members re-drawn with
 * replacement using pick_uniform[k * threshold] in [0, 1).  n <= 2048, k <= 16.
 * workspace >= scda_kmeans_workspace_bytes(n, k).  One CTA; nothing synchronises. */
size_t scda_kmeans_workspace_bytes(int n, int k);
int scda_kmeans_regions(const float *rois, int roi_stride, int n, int k, int first_center_id,
                        const double *uniforms, int n_local_trials, int max_iter, float tol,
                        const float *pick_uniform, int threshold, int *labels, float *centers,
                        int *counts, long long *index, void *workspace, size_t workspace_bytes,
                        cudaStream_t stream);

/* --- RoI max pooling on the NHWC bf16 feature map ------------------------ */
/* the operator of ROIPoolForwardLaucher / ROIPoolBackwardLaucher above (same RoI rounding,
 * bin windows and first-maximum rule, roi_pooling_kernel.cu:24-93,128-203) in the layout of
 * the tensor-core stages: feat [batch,H,W,C] bf16 -> out [num_rois, C, PH, PW] bf16 (the
 * reference's channel-major order = fc6's input) and argmax [num_rois, C, PH, PW] uint16 =
 * h * W + w inside the RoI's image, 0xFFFF for an empty bin.  C % 8 == 0, H * W < 65535.
 * Backward: dfeat [batch,H,W,C] fp32 is zeroed, then every dout element is added at its
 * argmax (fp32 red.add: sums are order-independent up to fp32 rounding). */
int scda_roi_pool_nhwc_bf16_fwd(const void *feat, float spatial_scale, int num_rois, int batch, int H,
                                int W, int C, int PH, int PW, const float *rois, void *out,
                                unsigned short *argmax, cudaStream_t stream);
int scda_roi_pool_nhwc_bf16_bwd(const void *dout, const unsigned short *argmax, const float *rois,
                                int num_rois, int batch, int H, int W, int C, int PH, int PW,
                                float *dfeat, cudaStream_t stream);

/* the same on an fp32 feature map / fp32 out and dout (the fp32-parity precision mode below) */
int scda_roi_pool_nhwc_f32_fwd(const float *feat, float spatial_scale, int num_rois, int batch, int H,
                               int W, int C, int PH, int PW, const float *rois, float *out,
                               unsigned short *argmax, cudaStream_t stream);
int scda_roi_pool_nhwc_f32_bwd(const float *dout, const unsigned short *argmax, const float *rois,
                               int num_rois, int batch, int H, int W, int C, int PH, int PW,
                               float *dfeat, cudaStream_t stream);

/* --- fp32-parity precision mode ("bf16x3") ------------------------------- */
/* The reference's convolutions / linear layers are fp32 (cuDNN / cuBLAS of torch 0.4.1:
 * vgg_adver_expansion_cluster.py:46-60,101-114, models/head.py:13-18, common_net.py).  In this
 * mode activations and gradients stay fp32 NHWC and the tensor-core entry points above are fed
 * operands split into bf16 halves (x = hi + lo) concatenated along the reduction dimension,
 * A = [hi|lo|hi], B = [hi|hi|lo]: three MMAs per product, ~2^-16 relative error per product.
 * The epilogue flag 64 (mask_src holds fp32) goes with it.
 * scda_split3_f32_bf16: x fp32 [rows, C] (row stride ldx) -> y bf16 [rows, 3C] = [hi|lo|hi]; C % 8 == 0.
 * scda_split_weights_f32_bf16: w fp32 [rows, K] (row stride ldw) -> fwd bf16 [rows, 3K] = [hi|hi|lo]
 * (K-major B operand of the forward GEMM / convolution) and / or stk bf16 [3 rows, K] = [hi;hi;lo]
 * (the MN-major B operand of the data gradient); either may be NULL.  K % 8 == 0. */
int scda_split3_f32_bf16(long long rows, int C, const float *x, long long ldx, void *y, cudaStream_t stream);
int scda_split_weights_f32_bf16(long long rows, int K, const float *w, long long ldw, void *fwd, void *stk,
                                cudaStream_t stream);
/* scda_conv3x3_wgrad_bf16_nhwc on channel sub-blocks of wider tensors: ldx / ldy = elements between
 * consecutive pixels (how the hi / lo halves of a split tensor are addressed) */
int scda_conv3x3_wgrad_bf16_nhwc_ld(int NB, int H, int W, int Cin, int Cout, const void *x, long long ldx,
                                    const void *dy, long long ldy, float *dw_partials, int splits,
                                    cudaStream_t stream);
/* fp32 NHWC forms of scda_maxpool2x2_*_bf16 / scda_nchw_f32_to_nhwc_bf16 (C % 4 == 0) */
int scda_maxpool2x2_nhwc_f32(int NB, int H, int W, int C, const float *x, float *y, cudaStream_t stream);
int scda_maxpool2x2_bwd_nhwc_f32(int NB, int H, int W, int C, const float *x, const float *dy, float *dx,
                                 int relu_mask, cudaStream_t stream);
int scda_nchw_f32_to_nhwc_f32(int NB, int C, int H, int W, int Cpad, const float *x, float *y,
                              cudaStream_t stream);
/* out[N] += column sums of x fp32 [M, ld] (any N, any ld >= N) */
int scda_colsum_f32_ld(long long M, int N, const float *x, long long ld, float *out, cudaStream_t stream);

/* --- discriminators (GAN_dis_AE, GAN_dis_AE_patch) ---------------------- */
/* replaces cuDNN / torch behind nn.Conv2d(k=3, s=2, p=1) (+ LeakyReLU / BatchNorm2d) and the 1x1 head of
 * faster_rcnn_adver_expansion_reweight_cluster.py:270-333, common_net.py:205-261.
 * Stride-2 3x3 convolution on the tcgen05 halo kernel: x bf16 NHWC [NB, 2Ho, 2Wo, C] (C % 32 == 0), wd bf16
 * [Cout][3][3][4C] = the phase-decomposed weight layout written by scda_conv_s2_weights from bf16
 * [Cout][3][3][C]; flags / slope: epilogue flags of the GEMM entry points (1 ReLU, 2 fp32 out, 4 mask,
 * 64 fp32 mask, 128 LeakyReLU(slope), 256 mask with LeakyReLU gradient). */
int scda_conv_s2_weights(int Cout, int C, const void *w_krsc_bf16, void *wd_bf16, cudaStream_t stream);
int scda_conv3x3_s2_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *x, const void *wd,
                              const float *bias, void *y, int flags, float slope, cudaStream_t stream);
/* dx [NB, 2Ho, 2Wo, C] from dy [NB, Ho, Wo, Cout]; mask_src (indexed like dx) = the input activation */
int scda_conv3x3_s2_dgrad_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *dy, const void *wd,
                                    void *dx, int flags, const void *mask_src, float slope, cudaStream_t stream);
/* dw_partials fp32 [splits][Cout][9][4C] (taps 0, 1, 3, 4 written); scda_conv_s2_wgrad_gather sums the
 * slabs and folds them to dw fp32 [Cout][3][3][C] (+= when accumulate) */
int scda_conv3x3_s2_wgrad_bf16_nhwc(int NB, int Ho, int Wo, int C, int Cout, const void *x, const void *dy,
                                    float *dw_partials, int splits, cudaStream_t stream);
int scda_conv_s2_wgrad_gather(int Cout, int C, const float *partials, int splits, float *dw, int accumulate,
                              cudaStream_t stream);
/* first layer Conv2d(3, 32, 3, s 2, p 1) + LeakyReLU, direct: x fp32 [N, 3, H, W] with element strides
 * (sn, sc, sh, sw); w fp32 [32][3][3][3] (o, r, s, c); y NHWC [N, H/2, W/2, 32] bf16 (y_f32 = 0) or fp32.
 * Backward: g = gradient w.r.t. the pre-activation; dw [32][3][3][3], db [32] (= or +=), dx fp32 NHWC
 * [N, H, W, 3]; any of dw / dx may be NULL.  workspace >= scda_disc_l1_workspace_bytes when dw != NULL. */
int scda_disc_l1_fwd(int N, int H, int W, const float *x, long long sn, long long sc, long long sh,
                     long long sw, const float *w_orsc, const float *bias, float slope, void *y, int y_f32,
                     cudaStream_t stream);
size_t scda_disc_l1_workspace_bytes(int N, int H, int W);
int scda_disc_l1_bwd(int N, int H, int W, const float *x, long long sn, long long sc, long long sh,
                     long long sw, const float *w_orsc, const void *g, int g_f32, float *dw, float *db,
                     float *dx, int accumulate, void *workspace, size_t workspace_bytes, cudaStream_t stream);
/* 1x1 head Conv2d(C, 1, 1): out[p] = bias + x[p, :] . w.  Backward: dx[p, c] = g[p] w[c] (x > 0 ? 1 : slope)
 * (through the LeakyReLU that produced x), dw[c] += sum_p g x, db += sum g (dw / db must be zeroed or hold
 * the running sum).  C divides 256. */
int scda_head_dot_fwd(long long P, int C, const void *x, int x_f32, const float *w, const float *bias,
                      float *out, cudaStream_t stream);
int scda_head_dot_bwd(long long P, int C, const void *x, int x_f32, const float *w, const float *g,
                      float slope, void *dx, float *dw, float *db, cudaStream_t stream);
/* nn.BatchNorm2d in training mode + LeakyReLU over x fp32 [P, C] (P = N*H*W): batch statistics, running
 * statistics updated (momentum, unbiased variance) when the pointers are not NULL; mean / rstd [C] saved for
 * the backward.  Backward: dgamma, dbeta (= or +=), dx (may be NULL).  Two launches each (per-chunk partial
 * sums, then finish + apply); workspace: scda_bn_workspace_bytes(P, C) bytes, contents irrelevant. */
size_t scda_bn_workspace_bytes(long long P, int C);
int scda_bn_lrelu_fwd(long long P, int C, const float *x, const float *gamma, const float *beta, float eps,
                      float slope, float momentum, float *running_mean, float *running_var, float *mean,
                      float *rstd, void *y, int y_f32, void *workspace, size_t workspace_bytes,
                      cudaStream_t stream);
int scda_bn_lrelu_bwd(long long P, int C, const float *x, const void *dy, int dy_f32, const float *gamma,
                      const float *beta, const float *mean, const float *rstd, float slope, void *dx,
                      int dx_f32, float *dgamma, float *dbeta, int accumulate, void *workspace,
                      size_t workspace_bytes, cudaStream_t stream);
/* global average pool x fp32 [N, HW, C] -> [N, C] and its backward (dx bf16 or fp32 [N, HW, C]) */
int scda_avgpool_fwd(int N, int HW, int C, const float *x, float *out, cudaStream_t stream);
int scda_avgpool_bwd(int N, int HW, int C, const float *g, void *dx, int dx_f32, cudaStream_t stream);

/* --- input pipeline ---------------------------------------------------- */
/* decoded uint8 HWC image [H0, W0, 3] (device) -> fp32 CHW [3, H, W]: resize (mode 0 nearest as Pillow < 7's
 * Image.resize, 1 bilinear), optional left-right mirror, / 255, (x - mean) / std — what
 * datasets/example_dataset.py:86-104,129-145 and target_dataset.py:41-71 do on the host with PIL + torchvision.
 * mean3 / std3: HOST pointers to three floats. */
int scda_image_prepare(const unsigned char *src_hwc, int H0, int W0, float *dst_chw, int H, int W, int mode,
                       int flip, const float *mean3, const float *std3, cudaStream_t stream);

/* --- detector losses -------------------------------------------------- */
/* F.cross_entropy(logits, targets, ignore_index) (mean over the counted rows) AND the reference's
 * top-1 `accuracy` in one pass (_add_rpn_loss / _add_rcnn_loss,
 * faster_rcnn_adver_expansion_reweight_cluster.py:36-68, 249-267).  logits fp32 [M, C] with row stride
 * ld, C <= 32; targets int64 [M].  out3 = {loss, accuracy in percent, rows counted} (device floats; loss
 * is NaN when no row is counted, as torch).  workspace >= scda_softmax_ce_workspace_bytes(M).
 * Backward: dlogits[m, c] = grad_loss[0] * (softmax - onehot) / rows counted, 0 for ignored rows. */
size_t scda_softmax_ce_workspace_bytes(long long M);
int scda_softmax_ce_acc_fwd(long long M, int C, const float *logits, long long ld, const long long *targets,
                            long long ignore_index, float *out3, void *workspace, size_t workspace_bytes,
                            cudaStream_t stream);
int scda_softmax_ce_bwd(long long M, int C, const float *logits, long long ld, const long long *targets,
                        long long ignore_index, const float *stats3, const float *grad_loss, float *dlogits,
                        long long lddx, cudaStream_t stream);
/* RPN objectness (…reweight_cluster.py:153-155, functions/rpn_proposal.py:44-49): 2-way softmax over each
 * anchor's (bg, fg) channel pair of the NCHW class map [B, 2A, H, W]; the foreground probability is written
 * in the proposal stage's anchor order, scores[b, (h*W + w)*A + a]. */
int scda_rpn_fg_scores(int B, int A, int H, int W, const float *cls_nchw, float *scores, cudaStream_t stream);

/* --- optimiser -------------------------------------------------------- */
/* replaces torch.optim.Adam(...).step() on each of the four networks
 * (tools/faster_rcnn_train_val.py:305-316 construct, :616,:635,:704,:750 step): one pass
 * over a network's FLAT fp32 parameter / gradient / moment buffers (torch 0.4.1 update
 * rule, weight decay added to the gradient).  grad_scale multiplies the gradient first
 * (1 when the loss was already divided by world size).  bf16_shadow, if not NULL, receives
 * the updated parameters rounded to bf16 for the tensor-core kernels.  step counts from 1.
 * lr_t_dev, if not NULL, is a device float holding the bias-corrected step size
 * lr * sqrt(1 - beta2^t) / (1 - beta1^t) for this step (then `step` and `lr` are ignored):
 * a captured CUDA graph replays the launch while the host only refreshes that scalar. */
int scda_adam_step(float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                   void *bf16_shadow, long long n, int step, float lr, float beta1, float beta2,
                   float eps, float weight_decay, float grad_scale, const float *lr_t_dev,
                   cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SCDA_B200_H_ */
