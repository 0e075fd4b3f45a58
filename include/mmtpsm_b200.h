/*
 * mmtpsm_b200.h — C ABI of libmmtpsm_b200.so
 *
 * The B200 (sm_100a) replacement for the native side of the MMT-PSM mean-teacher training hot
 * path.  Every entry point takes raw DEVICE pointers, sizes and a CUDA stream (passed as void*,
 * i.e. a cudaStream_t), returns an int status (0 = ok) and never allocates, frees or synchronises:
 * the caller owns every output and workspace buffer.  No torch types appear in this file.
 * The ONLY host-memory arguments are the small tables whose parameter names end in `_host` (segment offsets, per-level
 * pointer / size tables): they are read during the call and become kernel arguments.
 *
 * Process-wide state is limited to instrumentation and tuning knobs, none of which changes a result: the launch counter
 * and the profiling switch below (bench.py only), and the conv engine's mmt_conv_set_max_ctas / mmt_conv_force_cta_group /
 * mmt_conv_set_pdl / mmt_conv_set_pipeline.  Errors are reported through a thread-local string.
 *
 * Each block cites the reference interface it replaces (paths relative to the reference root).
 * The reference-side binding a maintainer adds is shown in INTEGRATION.md.
 */
#ifndef MMTPSM_B200_H_
#define MMTPSM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_OK 0
#define MMT_ERR_INVALID_ARGUMENT 1
#define MMT_ERR_WORKSPACE_TOO_SMALL 2
#define MMT_ERR_CUDA 3
#define MMT_ERR_UNSUPPORTED 4

/* Thread-local text of the last error raised by a call on this thread ("" if none). */
const char* mmt_last_error(void);
/* Library version string, e.g. "mmtpsm_b200 0.1 (sm_100a)". */
const char* mmt_version(void);
/* Number of kernels this library has launched on this process since load / last reset
 * (used by bench.py for the "gpu_launches" claim). */
int64_t mmt_launch_count(void);
void mmt_reset_launch_count(void);
/* Kernels of this library replayed from a CUDA graph are launched by the driver, not through the entry points that
 * count; the host wrapper that replays a captured graph adds the number it recorded at capture time. */
void mmt_add_launch_count(int64_t n);

/* Per-kernel-family device timing (off by default; used by bench.py's roofline leg only).
 * While enabled, every launch of a profiled family is bracketed by two cudaEvents recorded on the
 * launching stream; `mmt_prof_collect` synchronises those events and returns, per family,
 * the summed device time [ms], the ALGORITHMIC flops and bytes and the launch count.
 * Families: */
#define MMT_PROF_CONV_FPROP 0   /* conv_tc_kernel, forward (tcgen05 implicit GEMM)          */
#define MMT_PROF_CONV_DGRAD 1   /* conv_tc_kernel, data gradient                            */
#define MMT_PROF_CONV_WGRAD 2   /* wgrad_tc_kernel, weight gradient                         */
#define MMT_PROF_CONV_SIMT 3    /* fp32 FFMA convs (stem 7x7, odd channel counts)           */
#define MMT_PROF_ROI_ALIGN 4    /* ROIAlign forward + backward                              */
#define MMT_PROF_OPTIM 5        /* fused SGD + EMA                                          */
#define MMT_PROF_OTHER 6        /* every other kernel of this library                       */
#define MMT_PROF_FAMILIES 7
void mmt_prof_enable(int on);
int mmt_prof_collect(double* ms, double* flops, double* bytes, int64_t* launches, int n_families);

/* ------------------------------------------------------------------------------------------
 * NMS — replaces maskrcnn_benchmark/csrc/nms.h:10-28 (`_C.nms`) and csrc/cuda/nms.cu:23-131.
 * Semantics are those of the CUDA flavour (the one the reference's live path always takes,
 * structures/boxlist_ops.py:28-30): legacy "+1" areas (nms.cu:13-21), suppress when IoU > thr
 * (nms.cu:60), boxes visited in descending score order, kept ORIGINAL indices returned in
 * ascending order (nms.cu:127-130).  Ties in score are broken towards the lower original index.
 * Unlike the reference there is no device->host copy, no malloc and no legacy-stream launch.
 *
 * Segmented form: `n_segments` independent problems laid out back to back; segment s owns boxes
 * [seg_offsets[s], seg_offsets[s+1]).  `seg_offsets` is a HOST array of n_segments+1 ints.
 * keep_out   [total]      int64, per segment the kept indices RELATIVE to the segment start,
 *                         ascending, written at keep_out + seg_offsets[s]
 * n_keep_out [n_segments] int32 (device) number kept per segment
 * ------------------------------------------------------------------------------------------ */
size_t mmt_nms_workspace_bytes(int total_boxes, int n_segments, int max_segment_boxes);
int mmt_nms_segmented_f32(const float* boxes_xyxy, const float* scores, const int* seg_offsets_host,
                          int n_segments, float iou_threshold, int64_t* keep_out, int* n_keep_out,
                          void* workspace, size_t workspace_bytes, void* stream);
/* ------------------------------------------------------------------------------------------
 * Proposal decoding: BoxCoder.decode + BoxList.clip_to_image(remove_empty=False) + remove_small_boxes in one pass
 * (reference modeling/box_coder.py:56-95, structures/bounding_box.py:211-221, structures/boxlist_ops.py:37-54,
 * called per level and image from modeling/rpn/inference.py:107-124).  Candidates are laid out image-major:
 * `per_image` consecutive rows belong to image i = row / per_image.
 * deltas, anchors [n_images*per_image, 4] fp32 (16-byte aligned); scores [rows] or NULL;
 * image_wh [n_images, 2] fp32 (width, height) on the device, or NULL for no clipping;
 * weights (wx, wy, ww, wh), xform_clip = log(1000/16), min_size <= 0 disables the size filter.
 * boxes [rows,4]; scores_nms [rows] (= score if the box passes the size filter else -1) or NULL;
 * valid [rows] uint8 or NULL. */
int mmt_decode_clip_boxes_f32(const float* deltas, const float* anchors, const float* scores,
                              const float* image_wh, int n_images, int per_image, float wx, float wy,
                              float ww, float wh, float xform_clip, float min_size, float* boxes,
                              float* scores_nms, unsigned char* valid, void* stream);
/* ------------------------------------------------------------------------------------------
 * Fused boxlist_iou + Matcher (reference structures/boxlist_ops.py:57-92, modeling/matcher.py:44-138, top_k < 2):
 * matches_out[a] = index of the ground-truth box with the highest IoU (+1 extents) for prediction a, or -1 (best IoU
 * below low_threshold) / -2 (between the thresholds); with allow_low_quality_matches every prediction that attains
 * some ground truth's maximum IoU gets its arg-max back.  The [n_gt, n_boxes] IoU matrix is never materialised.
 * gt / boxes: xyxy fp32, 16-byte aligned; n_gt <= 1024; workspace >= mmt_match_workspace_bytes(). */
size_t mmt_match_workspace_bytes(int n_gt, int n_boxes);
int mmt_match_boxes_f32(const float* gt_xyxy, int n_gt, const float* boxes_xyxy, int n_boxes, float high_threshold,
                        float low_threshold, int allow_low_quality_matches, int64_t* matches_out, void* workspace,
                        size_t workspace_bytes, void* stream);
/* BoxCoder.encode fused with the gather of the matched ground truth (reference modeling/box_coder.py:22-54; call sites
 * rpn/loss.py:79-82, roi_heads/box_head/loss.py:75-78): targets[t] = encode(gt[max(matched_idx[t], 0)], boxes[t]);
 * matched_idx NULL = row-wise pairs.  16-byte aligned box arrays. */
int mmt_encode_boxes_f32(const float* gt_xyxy, const int64_t* matched_idx, const float* boxes_xyxy, int64_t n_boxes,
                         float wx, float wy, float ww, float wh, float* targets_out, void* stream);
/* ROI table + FPN level of every box of a pooler call in one launch (reference modeling/poolers.py:78-89
 * convert_to_roi_format and :31-42 LevelMapper): boxes = the per-image box lists concatenated, image_counts_host[i]
 * boxes belong to image i (<= 64 images).  rois_out [K,5] = (image index, x1, y1, x2, y2), levels_out [K] int32 =
 * clamp(floor(canonical_level + log2(sqrt(area) / canonical_scale + eps)), k_min, k_max) - k_min. */
int mmt_fpn_rois_f32(const float* boxes_xyxy, const int* image_counts_host, int n_images, float k_min, float k_max,
                     float canonical_scale, float canonical_level, float eps, float* rois_out, int* levels_out,
                     void* stream);
/* Balanced positive / negative sampling for a batch of images in one launch (reference
 * modeling/balanced_positive_negative_sampler.py:16-71): labels [total] fp32 (>= 1 positive, 0 negative, anything else
 * ignored), image i owns rows [image_offsets_dev[i], image_offsets_dev[i+1]) (device int32 array of n_images + 1).
 * Per image a uniformly random min(#pos, quota_pos) positives and min(#neg, batch_size_per_image - taken positives)
 * negatives are flagged in pos_mask_out / neg_mask_out (uint8 [total]).  Counter-based random keys from `seed`: the
 * same distribution as the reference's randperm(...)[:n] draws, a different stream (no host round trip). */
int mmt_balanced_sample_f32(const float* labels, const int* image_offsets_dev, int n_images, int batch_size_per_image,
                            int quota_pos, unsigned long long seed, unsigned char* pos_mask_out,
                            unsigned char* neg_mask_out, void* stream);
/* Same, additionally writing keep_flags_out [total] uint8: 1 for every kept row (original row order), 0 otherwise —
 * what a caller needs to select survivors with masks instead of index lists (no count read-back). */
int mmt_nms_segmented_flags_f32(const float* boxes_xyxy, const float* scores, const int* seg_offsets_host,
                                int n_segments, float iou_threshold, int64_t* keep_out, int* n_keep_out,
                                unsigned char* keep_flags_out, void* workspace, size_t workspace_bytes, void* stream);
/* RPN candidate gather for one pyramid level (reference modeling/rpn/inference.py:86-105): given the top-k anchor
 * indices topk_idx [n_images, k] (index = pixel * num_anchors + anchor) and their scores, fetch the deltas from the
 * fused head map fused_nhwc [n_images, pixels, padded_channels] (channels [A, 5A) = deltas of anchor 0..A-1), the
 * anchors from level_anchors [pixels * A, 4], and write deltas / anchors / scores at columns [col_offset, col_offset + k)
 * of the [n_images, k_total, ...] candidate tables. */
int mmt_rpn_gather_f32(const float* fused_nhwc, int padded_channels, int num_anchors, int64_t pixels,
                       const int64_t* topk_idx, const float* topk_vals, const float* level_anchors, int n_images,
                       int k, int k_total, int col_offset, float* deltas_out, float* anchors_out, float* scores_out,
                       void* stream);
/* RPN candidate selection for ALL (image, level) problems of a pass in one launch: per row of `level_scores[l]` ([n_images,
 * n_l] sigmoid objectness in (h, w, anchor) order) the k_l highest scores in descending order — reference
 * rpn/inference.py:92-94 `objectness.topk(pre_nms_top_n, dim=1, sorted=True)`; equal scores are ordered by ascending
 * anchor index — fused with mmt_rpn_gather_f32: deltas / anchors / scores of the selected candidates are written to
 * columns [sum_{m<l} k_m, +k_l) of the [n_images, k_total] tables.  The level tables are HOST arrays of device pointers /
 * sizes.  k_l <= 2048 (MMT_ERR_UNSUPPORTED otherwise: use torch.topk + mmt_rpn_gather_f32). */
int mmt_rpn_topk_gather_f32(const float* const* level_scores_host, const float* const* level_fused_host,
                            const float* const* level_anchors_host, const int* level_n_host, const int* level_k_host,
                            int n_levels, int n_images, int num_anchors, int padded_channels, int k_total,
                            float* deltas_out, float* anchors_out, float* scores_out, void* stream);
/* Single problem convenience form (exactly `_C.nms(dets, scores, thr)` minus the tensor plumbing). */
int mmt_nms_f32(const float* boxes_xyxy, const float* scores, int n, float iou_threshold,
                int64_t* keep_out, int* n_keep_out, void* workspace, size_t workspace_bytes,
                void* stream);

/* ------------------------------------------------------------------------------------------
 * ROIAlign — replaces csrc/ROIAlign.h:11-45 (`_C.roi_align_forward/backward`) and
 * csrc/cuda/ROIAlign_cuda.cu:65-122,178-254.  "aligned=False" Caffe2 semantics: no coordinate
 * rounding, ROI extent clamped to >= 1, sampling grid = sampling_ratio (or ceil(roi/pooled) when
 * <= 0), out-of-bounds rule y < -1 || y > H -> 0.
 *
 * Feature maps and outputs are addressed through explicit ELEMENT strides so the same kernel
 * serves NCHW (the reference's layout) and NHWC/channels_last (the layout this library's conv
 * engine produces; channel stride 1 gives 128-byte coalesced, float4-vectorised accesses).
 * rois [n_rois,5] = (batch_index as float, x1, y1, x2, y2), contiguous.
 * Backward ACCUMULATES into grad_input (caller zero-fills it once; several levels / poolers can
 * share one buffer).
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t n, c, h, w; /* element strides of the batch, channel, row and column axes */
} mmt_strides4;

int mmt_roi_align_forward_f32(const float* input, mmt_strides4 input_strides, int batch, int channels,
                              int height, int width, const float* rois, int n_rois,
                              float spatial_scale, int pooled_h, int pooled_w, int sampling_ratio,
                              float* output, mmt_strides4 output_strides, void* stream);
int mmt_roi_align_backward_f32(const float* grad_output, mmt_strides4 grad_output_strides,
                               const float* rois, int n_rois, float spatial_scale, int pooled_h,
                               int pooled_w, int sampling_ratio, float* grad_input,
                               mmt_strides4 grad_input_strides, int batch, int channels, int height,
                               int width, void* stream);

/* Fused FPN pooler — replaces the per-level loop of modeling/poolers.py:91-121 (`Pooler.forward`:
 * per level one `nonzero`, one ROIAlign launch and one index_put scatter).  All levels in ONE
 * launch: ROI n samples level roi_levels[n] (int32, DEVICE; computed by the caller with the
 * reference's LevelMapper formula, poolers.py:31-42) and writes row n of the dense NHWC output
 * [n_rois, pooled_h, pooled_w, channels].  Level feature maps are dense NHWC [batch, h_l, w_l,
 * channels]; the pointer / size / scale tables are HOST arrays of n_levels (<= 8) entries.
 * Backward ACCUMULATES into the per-level gradient maps (caller zero-fills them). */
int mmt_roi_align_fpn_forward_f32(const float* const* level_feats_host, const int* level_h_host,
                                  const int* level_w_host, const float* level_scale_host,
                                  int n_levels, int batch, int channels, const float* rois,
                                  int n_rois, int pooled_h, int pooled_w, int sampling_ratio,
                                  const int* roi_levels, float* output, void* stream);
int mmt_roi_align_fpn_backward_f32(const float* grad_output, const float* rois,
                                   const int* roi_levels, int n_rois, int pooled_h, int pooled_w,
                                   int sampling_ratio, float* const* level_grads_host,
                                   const int* level_h_host, const int* level_w_host,
                                   const float* level_scale_host, int n_levels, int batch,
                                   int channels, void* stream);

/* ROIPool — replaces csrc/ROIPool.h:10-45 / csrc/cuda/ROIPool_cuda.cu:17-108 (kept for ABI
 * completeness; the FPN configuration on the hot path only uses ROIAlign). NCHW contiguous. */
int mmt_roi_pool_forward_f32(const float* input, int batch, int channels, int height, int width,
                             const float* rois, int n_rois, float spatial_scale, int pooled_h,
                             int pooled_w, float* output, int* argmax, void* stream);
int mmt_roi_pool_backward_f32(const float* grad_output, const int* argmax, const float* rois,
                              int n_rois, int pooled_h, int pooled_w, float* grad_input, int batch,
                              int channels, int height, int width, void* stream);

/* ------------------------------------------------------------------------------------------
 * Colour views of unlabeled crops (reference data/transforms/build.py:27-35 `transpart2`,
 * transforms.py:84-206, data/datasets/Pap.py:818-831): AdjustBrightness -> AdjustContrast -> AdjustHue -> RandomErasing ->
 * ToTensor -> Normalize(BGR255 - mean), bit-exact with the Pillow / torchvision arithmetic of the host pipeline.
 * images_hwc [batch, H, W, 3] uint8 RGB (device); per (view, image) row-major [n_views * batch]: brightness / contrast
 * factors (fp32), hue_shift (int: uint8 wrap-around amount, trunc(hue_factor * 255) mod 256).
 * mmt_color_luma_sums_u8: luma_sums_out[v*batch+b] = sum over pixels of L of the brightness-adjusted image (the contrast
 * pivot is int(sum / pixels + 0.5)).  mmt_color_views_u8: views_out [n_views, batch, 3, H, W] fp32 (channels B, G, R);
 * erase_rects = device array of n_rects records {int view*batch+image, top, left, h, w; int64 fill_offset} applied in
 * order, fill_pool = uint8 [h*w*3] blocks (the host draws rectangles and noise: the reference's random stream). */
int mmt_color_luma_sums_u8(const unsigned char* images_hwc, const float* brightness, int n_views, int batch, int height,
                           int width, unsigned long long* luma_sums_out, void* stream);
int mmt_color_views_u8(const unsigned char* images_hwc, const float* brightness, const float* contrast,
                       const int* hue_shift, const unsigned long long* luma_sums, int n_views, int batch, int height,
                       int width, const void* erase_rects, int n_rects, const unsigned char* fill_pool, float mean_b,
                       float mean_g, float mean_r, float* views_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * RPN loss (reference modeling/rpn/loss.py:138-196) over the fused head maps of all pyramid levels, in place:
 * fused_levels[l] = [n_images, level_pixels[l], padded_channels] fp32, channels [0, A) objectness logits of the A
 * anchors of a pixel, [A, 5A) their box deltas (host array of device pointers).  labels [n_images * sum(pixels*A)]
 * fp32 (image-major, level-major, (pixel, anchor) order; 1 / 0 / -1), regression_targets [same, 4], pos_mask /
 * neg_mask uint8 = the sampler's choice.  forward: sums_out[0] = sum of BCE-with-logits over the sampled anchors,
 * sums_out[1] = sum of smooth-L1(beta) over the positive anchors' deltas (the caller divides by |sampled|).
 * backward: writes d/d(fused) of  g_obj * sums[0] / n_sampled + g_box * sums[1] / n_sampled  into dfused_levels
 * (every element, zeros included); n_sampled, grad_objectness_loss, grad_box_loss are DEVICE scalars. */
int mmt_rpn_loss_forward_f32(const float* const* fused_levels, const int64_t* level_pixels, int n_levels,
                             int num_anchors, int padded_channels, int n_images, const float* labels,
                             const float* regression_targets, const unsigned char* pos_mask,
                             const unsigned char* neg_mask, float beta, float* sums_out, void* stream);
int mmt_rpn_loss_backward_f32(const float* const* fused_levels, float* const* dfused_levels,
                              const int64_t* level_pixels, int n_levels, int num_anchors, int padded_channels,
                              int n_images, const float* labels, const float* regression_targets,
                              const unsigned char* pos_mask, const unsigned char* neg_mask, float beta,
                              const float* n_sampled, const float* grad_objectness_loss, const float* grad_box_loss,
                              void* stream);

/* ------------------------------------------------------------------------------------------
 * SigmoidFocalLoss — ABSENT from this reference tree; semantics are those of
 * upstream facebookresearch/maskrcnn-benchmark csrc/cuda/SigmoidFocalLoss_cuda.cu:
 *   p = sigmoid(x); t = targets[n] (1-based class, 0 = background, < 0 = ignore)
 *   L[n,c] = -[t==c+1] a (1-p)^g log(max(p,FLT_MIN))
 *            -[t!=c+1 && t>=0] (1-a) p^g (-x [x>=0] - log(1+exp(x-2x[x>=0])))
 * losses/grad_logits are [n, num_classes]; backward multiplies by grad_losses[n,c].
 * ------------------------------------------------------------------------------------------ */
int mmt_sigmoid_focal_loss_forward_f32(const float* logits, const int* targets, int n,
                                       int num_classes, float gamma, float alpha, float* losses,
                                       void* stream);
int mmt_sigmoid_focal_loss_backward_f32(const float* logits, const int* targets,
                                        const float* grad_losses, int n, int num_classes,
                                        float gamma, float alpha, float* grad_logits, void* stream);

/* ------------------------------------------------------------------------------------------
 * PSM consistency loss — replaces the ~40-kernel chain of
 * modeling/roi_heads/box_head/loss.py:164-237 (`_mean_var_logits`, `evaluatePSM`),
 * :267-287 (`cls_loss`, 'bce' -> soft-target CE branch) and :311-315 (`sharpen`).
 *   student_logits [n,C], teacher_logits [n_views][n,C] (view-major, contiguous), labels [n] int64
 *   positives = labels > 0, negatives = labels == 0 (labels < 0 ignored)
 *   v[i]  = sum_c std_k(softmax(z_k[i]))[c]   (unbiased std over views)
 *   keep m = min(#neg, floor(#pos/2)) negatives of largest v (ties: lower index first);
 *            hard_negative == 0 keeps the first m negatives of the order given in `neg_order`
 *   q = softmax(mean_k z_k), sharpened q^(1/T)/sum when sharpen != 0
 *   loss = sum_c mean_rows(-q log_softmax(s) w) / 3, w = 1 (pos) / neg_weight (kept neg)
 * (the literal 3 is the reference's, loss.py:284,287, independent of C).
 * Outputs: loss_out[1]; grad_student [n,C] = d loss / d student_logits (zero on unused rows);
 * selected_out [n] int32 (1 = positive row, 2 = kept negative, 0 = unused) for exact-index parity.
 * Workspace: mmt_psm_workspace_bytes(n).
 * ------------------------------------------------------------------------------------------ */
size_t mmt_psm_workspace_bytes(int n);
int mmt_psm_loss_f32(const float* student_logits, const float* teacher_logits, int n_views,
                     const int64_t* labels, int n, int num_classes, float temperature, int sharpen,
                     int hard_negative, float neg_weight, const int* neg_order, float* loss_out,
                     float* grad_student, int* selected_out, void* workspace,
                     size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * MGD (mask-guided feature distillation) — replaces `fg_hint_loss`
 * (modeling/detector/generalized_rcnn.py:243-282), ~160 launches per iteration.
 *
 * mmt_mgd_masks: integral foreground map M [batch, H, W] (float) -> per-level binary masks
 *   m_l = 1[adaptive_avg_pool2d(M, (h_l, w_l)) > 0.5] (generalized_rcnn.py:259-264) and their
 *   sums.  masks_out is the concatenation over levels of [batch, h_l, w_l] uint8;
 *   mask_sums_out [n_levels] float.
 * mmt_mgd_level: one pyramid level, all teacher views in one pass (student read once):
 *   num[k] = sum_{b,y,x,c} (S - T_k')^2 m  with T_k' = T_k flipped along x when flip_mask bit k
 *   is set (generalized_rcnn.py:246-249).  Features are NHWC-addressed through strides.
 *   partial_out [n_views] float is ACCUMULATED (caller zero-fills).
 *   grad_student (nullable) receives sum_k coeff[k] * 2 (S - T_k') m, coeff = term_weight /
 *   (mask_sum * C + 1e-7) read from device (so no host sync is needed).
 * mmt_mgd_finalize: loss = mean over (view, level) of num / (mask_sum * C + 1e-7)
 *   (generalized_rcnn.py:267-281).
 * ------------------------------------------------------------------------------------------ */
int mmt_mgd_masks_f32(const float* integral_mask, int batch, int height, int width,
                      const int* level_h_host, const int* level_w_host, int n_levels,
                      uint8_t* masks_out, float* mask_sums_out, void* stream);
int mmt_mgd_level_f32(const float* student, mmt_strides4 student_strides,
                      const float* const* teacher_views_host, mmt_strides4 teacher_strides,
                      int n_views, unsigned flip_mask, const uint8_t* level_mask,
                      const float* level_mask_sum, int batch, int channels, int height, int width,
                      float term_weight, float* partial_out, float* grad_student,
                      mmt_strides4 grad_strides, void* stream);
int mmt_mgd_finalize_f32(const float* partials /*[n_levels][n_views]*/,
                         const float* mask_sums /*[n_levels]*/, int n_levels, int n_views,
                         int channels, float* loss_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mask-target rasterisation — replaces the per-positive-ROI HOST loop `project_masks_on_boxes`
 * (modeling/roi_heads/mask_head/loss.py:37-75): Polygons.crop / .resize / .convert("mask")
 * (structures/segmentation_mask.py:94-137) -> pycocotools rleFrPoly + rleMerge + rleDecode
 * (pycoco/maskApi.c:166-206, :53-75, :47-51).  Bit-exact (integer/double arithmetic restated).
 *   poly_xy            flat (x,y) fp32 vertices of every polygon of the image, image coordinates
 *   poly_offsets       [n_polygons+1] int32, vertex offsets into poly_xy/2
 *   inst_poly_offsets  [n_instances+1] int32, polygon ranges of each ground-truth instance
 *   roi_boxes [n_rois,4] xyxy fp32 (the positive proposals), roi_instance [n_rois] int32 (matched instance)
 *   masks_out [n_rois, mask_size, mask_size] fp32 in {0,1}.  All pointers are DEVICE pointers.
 * ------------------------------------------------------------------------------------------ */
int mmt_rasterize_polygons_f32(const float* poly_xy, const int* poly_offsets,
                               const int* inst_poly_offsets, const float* roi_boxes,
                               const int* roi_instance, int n_rois, int mask_size, float* masks_out,
                               void* stream);

/* ------------------------------------------------------------------------------------------
 * Teacher mask paste-and-sum — replaces `Masker.forward_single_image` / `paste_mask_in_image`
 * (modeling/roi_heads/mask_head/inference.py:169-231; one CPU canvas per detection) followed by
 * `.sum(0)` (modeling/detector/generalized_rcnn.py:129-132): integral_out[b,y,x] = number of
 * detections of image b whose pasted (padding 1, bilinear, > threshold) mask covers pixel (y,x).
 *   mask_probs [n_dets, M, M] fp32, boxes [n_dets,4] xyxy, image_index [n_dets] int32 (DEVICE)
 *   integral_out [n_images, im_h, im_w] fp32 — overwritten.
 * ------------------------------------------------------------------------------------------ */
int mmt_paste_masks_sum_f32(const float* mask_probs, const float* boxes, const int* image_index,
                            int n_dets, int mask_size, int n_images, int im_h, int im_w,
                            float threshold, float* integral_out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Teacher EMA + student SGD — replaces `MTtrainer.update_teacher`
 * (engine/MTtrainer.py:277-281; 264 launches) and torch.optim.SGD as configured by
 * solver/build.py:5-23 (one param group per tensor; bias lr x2, weight decay 0).
 * Parameters live in flat fp32 arenas; `segments` (DEVICE array) gives per-tensor hyper-params.
 *   SGD:  g = grad + wd * p;  buf = first_step ? g : momentum * buf + g;  p -= lr * buf
 *   EMA:  t = alpha * t + (1 - alpha) * p      (with the updated p, as the reference orders it)
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int64_t begin, end; /* element range in the arena; begin must be a multiple of 4 */
  float lr, weight_decay;
  int trainable; /* 0: SGD skipped (frozen tensor), EMA still applied */
  int reserved;
} mmt_param_segment;

int mmt_ema_f32(float* teacher, const float* student, int64_t n, float alpha, void* stream);
int mmt_sgd_ema_f32(float* student, float* teacher /*nullable: no EMA*/, const float* grad,
                    float* momentum_buf, int64_t n, const mmt_param_segment* segments_dev,
                    int n_segments, float momentum, float alpha, int first_step,
                    float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------
 * Convolution engine — replaces torch.nn.Conv2d + FrozenBatchNorm2d + ReLU as used by
 * layers/misc.py:30-43, layers/batch_norm.py:6-24, modeling/backbone/resnet.py:254-275,
 * modeling/backbone/fpn.py:43-69, modeling/rpn/rpn.py:39-46 (cuDNN in the reference).
 * Activations are NHWC fp32; weights are [out_c][k_h][k_w][in_c] ("OHWI") fp32.  The tensor
 * cores consume the fp32 operands as TF32 (kind::tf32, fp32 accumulate in TMEM) — the same
 * arithmetic cuDNN applies to the reference's fp32 convs on Ampere-and-later GPUs.
 *   y = act( conv(x, w) * scale[c] + bias[c] + residual )
 * `impl` selects the kernel: MMT_CONV_TCGEN05 (TMA + tcgen05 implicit GEMM; needs in_c % 32 == 0
 * and out_c % 16 == 0) or MMT_CONV_SIMT (plain fp32 FFMA kernel, any shape; used for the 7x7
 * stem (in_c = 3), the skinny heads, and as the on-device cross-check of the tensor-core path).
 * MMT_CONV_AUTO picks tcgen05 whenever the shape allows.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
  int batch, in_h, in_w, in_c;
  int out_c, k_h, k_w;
  int stride, pad;
  int out_h, out_w;
} mmt_conv_shape;

#define MMT_CONV_AUTO 0
#define MMT_CONV_TCGEN05 1
#define MMT_CONV_SIMT 2

#define MMT_EPI_RELU 1
#define MMT_STEM_INPUT_NCHW 2 /* mmt_conv_stem_fprop_f32 only: x is the reference's NCHW image batch [B,3,H,W] */

/* Persistent-grid size of the tcgen05 conv kernels launched (or captured into a CUDA graph) from now on: n CTAs,
 * one per SM, leaving 148 - n SMs free for kernels of other streams.  Default 148. */
int mmt_conv_set_max_ctas(int n);
/* Tile configuration of the tcgen05 fprop / dgrad kernel: 0 = the library's cost model chooses per problem between one
 * CTA per 128 x BN tile and a CTA pair (tcgen05 cta_group::2, cluster of 2) per 256 x BN tile; 1 / 2 force one of them
 * (parity tests and per-layer benchmarks). */
int mmt_conv_force_cta_group(int cta_group);
/* Programmatic dependent launch of the tcgen05 kernels (default on): their prologue (barrier init, TMEM allocation,
 * descriptor prefetch) overlaps the tail of the previous kernel of the stream; 0 = plain stream serialisation. */
int mmt_conv_set_pdl(int on);
/* Operand pipeline of the tcgen05 fprop / dgrad kernel: `stages` caps the shared-memory ring depth (0 = as deep as shared
 * memory allows: the default), `prefetch_kblocks` = how many 32-channel K blocks ahead of its loads the TMA producer issues
 * L2 prefetches of the activation tiles (0 = off). */
int mmt_conv_set_pipeline(int stages, int prefetch_kblocks);
size_t mmt_conv_workspace_bytes(const mmt_conv_shape* shape);
int mmt_conv_fprop_f32(const float* x_nhwc, const float* w_ohwi, const float* scale /*nullable*/,
                       const float* bias /*nullable*/, const float* residual_nhwc /*nullable*/,
                       float* y_nhwc, const mmt_conv_shape* shape, int epilogue_flags, int impl,
                       void* stream);
/* dx = conv_transpose(dy, w): data gradient.  dx is fully overwritten. */
/* ResNet stem (reference modeling/backbone/resnet.py:288-293: Conv2d(3, 64, 7, stride 2, pad 3) + FrozenBN + ReLU):
 * the 7x7 / stride-2 convolution over 3 channels runs on the tensor cores as a 4x4 implicit GEMM over the 2x2
 * space-to-depth image (12 channels padded to one 32-wide K block).  x NHWC [B,H,W,3] (or, with flag
 * MMT_STEM_INPUT_NCHW, the reference's NCHW [B,3,H,W] batch as is), w OHWI [out_c,7,7,3];
 * workspace (256-byte aligned) >= mmt_conv_stem_workspace_bytes() holds the space-to-depth image and filter. */
size_t mmt_conv_stem_workspace_bytes(const mmt_conv_shape* shape);
int mmt_conv_stem_fprop_f32(const float* x_nhwc, const float* w_ohwi, const float* scale /*nullable*/,
                            const float* bias /*nullable*/, float* y_nhwc, const mmt_conv_shape* shape, int flags,
                            void* workspace, size_t workspace_bytes, void* stream);
int mmt_conv_dgrad_f32(const float* dy_nhwc, const float* w_ohwi, float* dx_nhwc,
                       const mmt_conv_shape* shape, int impl, void* workspace,
                       size_t workspace_bytes, void* stream);
/* dw[o][r][s][i] = sum_pixels dy[p][o] x[p + (r,s)][i]; dbias[o] = sum dy (nullable).
 * dw/dbias are fully overwritten. */
int mmt_conv_wgrad_f32(const float* x_nhwc, const float* dy_nhwc, float* dw_ohwi,
                       float* dbias /*nullable*/, const mmt_conv_shape* shape, int impl,
                       void* workspace, size_t workspace_bytes, void* stream);
/* dw += wgrad(x, dy) into a live gradient buffer (tcgen05 path only; MMT_ERR_UNSUPPORTED otherwise): no memset, no
 * temporary, no separate accumulation pass. */
int mmt_conv_wgrad_accumulate_f32(const float* x_nhwc, const float* dy_nhwc, float* dw_ohwi, const mmt_conv_shape* shape,
                                  void* stream);

/* Fused backward of Conv2d -> FrozenBatchNorm2d [-> ReLU] chains (layers/batch_norm.py:19-24, resnet.py:254-275 under
 * autograd).  `dy` is the gradient of the BN OUTPUT; `k_scale` (nullable, per out channel) is the frozen-BN scale, folded
 * into the transformed filter; the epilogue combines the result with `aux` (dense NHWC, shaped like dx):
 *   MMT_DGRAD_AUX_ADD  : dx = conv_T(dy * scale, w) + aux        (gradient arriving through the skip connection)
 *   MMT_DGRAD_AUX_MASK : dx = aux > 0 ? conv_T(dy * scale, w) : 0  (ReLU backward of the layer whose OUTPUT aux is — this
 *                        conv's input)
 * tcgen05 path only (MMT_ERR_UNSUPPORTED otherwise; aux needs stride 1); workspace as mmt_conv_dgrad_f32. */
#define MMT_DGRAD_AUX_NONE 0
#define MMT_DGRAD_AUX_ADD 1
#define MMT_DGRAD_AUX_MASK 2
int mmt_conv_dgrad_fused_f32(const float* dy_nhwc, const float* w_ohwi, const float* k_scale /*nullable*/,
                             const float* aux_nhwc /*nullable*/, int aux_mode, float* dx_nhwc, const mmt_conv_shape* shape,
                             void* workspace, size_t workspace_bytes, void* stream);
/* dw (+)= row_scale[o] * wgrad(x, dy)[o]  (row_scale nullable; accumulate != 0: add into a live gradient buffer). */
int mmt_conv_wgrad_scaled_f32(const float* x_nhwc, const float* dy_nhwc, const float* row_scale /*nullable*/,
                              float* dw_ohwi, const mmt_conv_shape* shape, int accumulate, void* stream);

/* Fused elementwise helpers around the convs (all NHWC fp32, contiguous) ------------------- */
/* g_out = g_in * (y > 0) * scale[c]  (ReLU + FrozenBN backward; scale nullable, relu optional) */
int mmt_bn_relu_backward_f32(const float* grad_y, const float* y, const float* scale, float* grad_z,
                             int64_t pixels, int channels, int apply_relu_mask, void* stream);
/* 3x3/s2/p1 max-pool forward + backward (resnet.py:292 stem) and 1x1/s2 subsample (fpn.py:74). */
int mmt_maxpool3x3s2_forward_f32(const float* x, float* y, int batch, int in_h, int in_w,
                                 int channels, void* stream);
int mmt_subsample2_forward_f32(const float* x, float* y, int batch, int in_h, int in_w,
                               int channels, void* stream);
/* grad_x[2oy][2ox] += grad_y[oy][ox] (grad_x is accumulated into, in_h/in_w are x's extent). */
int mmt_subsample2_backward_f32(const float* grad_y, float* grad_x, int batch, int in_h, int in_w,
                                int channels, void* stream);
/* FPN top-down: out = lateral + nearest_upsample_x2(top)  (fpn.py:57-62), and its backward
 * (grad_top[y/2][x/2] += grad_out[y][x]).  lateral/out [b,h,w,c], top [b,th,tw,c]. */
int mmt_upsample2_add_f32(const float* lateral, const float* top, float* out, int batch, int h,
                          int w, int top_h, int top_w, int channels, void* stream);
int mmt_upsample2_backward_f32(const float* grad_out, float* grad_top, int batch, int h, int w,
                               int top_h, int top_w, int channels, void* stream);
/* Layout transforms between the reference's NCHW/OIHW and this library's NHWC/OHWI. */
int mmt_nchw_to_nhwc_f32(const float* src, float* dst, int n, int c, int h, int w, void* stream);
int mmt_nhwc_to_nchw_f32(const float* src, float* dst, int n, int c, int h, int w, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMTPSM_B200_H_ */
