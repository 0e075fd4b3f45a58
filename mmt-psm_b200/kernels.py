"""Tensor-level entry points over the C ABI (raw pointers in, caller-owned outputs).

Every function takes CUDA fp32 tensors, allocates outputs / workspace through torch's caching
allocator, launches on torch's CURRENT stream and never synchronises unless the reference API it
mirrors forces it (only `nms`, which must return a tensor whose length is data dependent).
There is no CPU path: non-CUDA tensors raise.
"""
import ctypes
from ctypes import c_float, c_int, c_void_p

import torch

from . import lib as _l
from .lib import ConvShape, Strides4, check

CONV_AUTO, CONV_TCGEN05, CONV_SIMT = 0, 1, 2
EPI_RELU = 1

# Process-wide conv math mode, the analogue of `torch.backends.cudnn.allow_tf32` (which the reference leaves at its
# default, True): "tf32" = tcgen05 kind::tf32 implicit GEMM wherever the shape allows (default), "fp32" = the fp32
# FFMA kernels everywhere (bit-for-bit fp32 multiply-adds; used by the step-level parity tests).
_conv_impl = CONV_AUTO


def set_conv_math(mode):
    global _conv_impl
    if mode not in ("tf32", "fp32"):
        raise ValueError("conv math mode must be 'tf32' or 'fp32'")
    _conv_impl = CONV_AUTO if mode == "tf32" else CONV_SIMT


def get_conv_math():
    return "tf32" if _conv_impl == CONV_AUTO else "fp32"


_conv_max_ctas = None   # None = the library's default (one CTA per SM)


def set_conv_max_ctas(n):
    """Size of the conv engine's persistent grids from now on (None = all 148 SMs); returns the previous setting.  A CUDA
    graph captured under a smaller setting keeps it: the trainer captures the TEACHER's trunk with a few SMs left free, so
    that the small head / glue kernels of the other pass never queue behind a full wave of persistent conv CTAs."""
    global _conv_max_ctas
    prev = _conv_max_ctas
    check(_l.lib().mmt_conv_set_max_ctas(int(n) if n else 148), "conv_set_max_ctas")
    _conv_max_ctas = int(n) if n else None
    return prev


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("mmtpsm_b200 ops run on CUDA tensors only (no CPU fallback); got %s" % t.device)


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw accessor is ~20x cheaper than building a
    torch.cuda.Stream object per launch: this runs before every kernel of the library)."""
    if _raw_stream is not None:
        return c_void_p(_raw_stream(torch.cuda.current_device()))
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _strides(t, order):
    """Element strides of a 4-D tensor as (n, c, h, w) given which dims those are."""
    st = t.stride()
    return Strides4(st[order[0]], st[order[1]], st[order[2]], st[order[3]])


# --------------------------------------------------------------------------------------- NMS
def nms_segmented(boxes, scores, seg_offsets, iou_threshold):
    """Independent NMS problems laid out back to back, one launch chain, no host sync.

    boxes [T,4] xyxy, scores [T]; seg_offsets: python list of S+1 ints.
    Returns (keep [T] int64 — per segment the kept indices relative to the segment start, ascending,
    stored from seg_offsets[s]; counts [S] int32), both on the device.
    """
    _need_cuda(boxes, scores)
    L = _l.lib()
    boxes = _f32c(boxes)
    scores = _f32c(scores)
    S = len(seg_offsets) - 1
    total = int(seg_offsets[-1])
    keep = torch.empty(max(total, 1), dtype=torch.int64, device=boxes.device)
    counts = torch.zeros(max(S, 1), dtype=torch.int32, device=boxes.device)
    if S == 0 or total == 0:
        return keep[:total], counts[:S]
    max_n = max(seg_offsets[i + 1] - seg_offsets[i] for i in range(S))
    ws_bytes = L.mmt_nms_workspace_bytes(total, S, max_n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=boxes.device)
    offs = (c_int * (S + 1))(*[int(o) for o in seg_offsets])
    check(L.mmt_nms_segmented_f32(_p(boxes), _p(scores), offs, S, float(iou_threshold), _p(keep),
                                  _p(counts), _p(ws), ws_bytes, _stream()), "nms")
    return keep[:total], counts[:S]


def nms_segmented_flags(boxes, scores, seg_offsets, iou_threshold):
    """As nms_segmented, returning only the per-row keep flags (bool [T], original row order): survivors are selected
    with masks, nothing is read back."""
    _need_cuda(boxes, scores)
    L = _l.lib()
    boxes, scores = _f32c(boxes), _f32c(scores)
    S = len(seg_offsets) - 1
    total = int(seg_offsets[-1])
    flags = torch.empty(total, dtype=torch.bool, device=boxes.device)
    if S == 0 or total == 0:
        return flags
    keep = torch.empty(total, dtype=torch.int64, device=boxes.device)
    counts = torch.empty(S, dtype=torch.int32, device=boxes.device)
    max_n = max(seg_offsets[i + 1] - seg_offsets[i] for i in range(S))
    ws_bytes = L.mmt_nms_workspace_bytes(total, S, max_n)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=boxes.device)
    offs = seg_offsets if isinstance(seg_offsets, ctypes.Array) else (c_int * (S + 1))(*[int(o) for o in seg_offsets])
    check(L.mmt_nms_segmented_flags_f32(_p(boxes), _p(scores), offs, S, float(iou_threshold), _p(keep), _p(counts),
                                        _p(flags), _p(ws), ws_bytes, _stream()), "nms_flags")
    return flags


def rpn_gather(fused, num_anchors, topk_idx, topk_vals, level_anchors, deltas_out, anchors_out, scores_out, col_offset):
    """Candidates of one pyramid level into the pass-wide tables (see mmt_rpn_gather_f32)."""
    L = _l.lib()
    N, k = topk_idx.shape
    assert fused.is_contiguous() and topk_idx.is_contiguous() and topk_vals.is_contiguous()
    pixels = fused.shape[1] * fused.shape[2]
    check(L.mmt_rpn_gather_f32(_p(fused), fused.shape[3], num_anchors, pixels, _p(topk_idx), _p(topk_vals),
                               _p(level_anchors), N, k, deltas_out.shape[1], col_offset, _p(deltas_out), _p(anchors_out),
                               _p(scores_out), _stream()), "rpn_gather")


def rpn_topk_gather(scores, fused, num_anchors, level_anchors, ks, deltas_out, anchors_out, scores_out):
    """All levels of a pass in ONE launch: per image and level the ks[l] best objectness scores (descending; ties by
    ascending anchor index) and, for them, the deltas (from the fused head maps), anchors and scores written into the
    pass-wide [N, sum(ks)] tables.  Returns False — nothing done — if a level asks for more than the kernel's 2048."""
    _need_cuda(*scores, *fused, *level_anchors, deltas_out, anchors_out, scores_out)
    L = _l.lib()
    n_levels, N = len(scores), scores[0].shape[0]
    sc = [_f32c(t) for t in scores]
    fu = [_f32c(t) for t in fused]
    an = [_f32c(t) for t in level_anchors]
    arr = lambda ts: (ctypes.c_void_p * n_levels)(*[t.data_ptr() for t in ts])
    ns = (ctypes.c_int * n_levels)(*[int(t.shape[1]) for t in sc])
    kk = (ctypes.c_int * n_levels)(*[int(k) for k in ks])
    rc = L.mmt_rpn_topk_gather_f32(arr(sc), arr(fu), arr(an), ns, kk, n_levels, N, int(num_anchors), int(fu[0].shape[-1]),
                                   int(sum(ks)), _p(deltas_out), _p(anchors_out), _p(scores_out), _stream())
    if rc == 4:
        return False
    check(rc, "rpn_topk_gather")
    return True


def nms(dets, scores, iou_threshold):
    """`_C.nms(dets, scores, thr)` (reference csrc/nms.h:10-28): int64 kept indices, ascending.
    Empty input returns an empty int64 CPU tensor exactly like the reference (nms.h:17-18)."""
    if dets.numel() == 0:
        return torch.empty(0, dtype=torch.int64, device="cpu")
    keep, counts = nms_segmented(dets, scores, [0, dets.shape[0]], iou_threshold)
    return keep[: int(counts[0].item())]


def decode_clip_boxes(deltas, anchors, weights, xform_clip, image_wh=None, per_image=None, scores=None, min_size=0.0):
    """Fused BoxCoder.decode + clip_to_image + remove_small_boxes over image-major candidate rows.
    deltas / anchors [T,4]; image_wh device tensor [N,2] (w,h) or None; returns (boxes [T,4], scores_nms [T] or None,
    valid [T] bool or None)."""
    _need_cuda(deltas, anchors, scores, image_wh)
    L = _l.lib()
    deltas, anchors = _f32c(deltas), _f32c(anchors)
    T = deltas.shape[0]
    assert anchors.shape == deltas.shape == (T, 4)
    boxes = torch.empty_like(deltas)
    want_valid = min_size > 0
    s_nms = torch.empty(T, dtype=torch.float32, device=deltas.device) if scores is not None else None
    valid = torch.empty(T, dtype=torch.bool, device=deltas.device) if want_valid else None
    if image_wh is not None:
        n_images = image_wh.shape[0]
        per_image = T // max(n_images, 1) if per_image is None else per_image
        assert n_images * per_image == T
    else:
        n_images, per_image = 1, T
    if scores is not None:
        scores = _f32c(scores)
    wx, wy, ww, wh = weights
    check(L.mmt_decode_clip_boxes_f32(_p(deltas), _p(anchors), _p(scores), _p(image_wh), n_images, per_image, float(wx),
                                      float(wy), float(ww), float(wh), float(xform_clip), float(min_size), _p(boxes),
                                      _p(s_nms), _p(valid), _stream()), "decode_clip_boxes")
    return boxes, s_nms, valid


def match_boxes(gt, boxes, high_threshold, low_threshold, allow_low_quality_matches):
    """Fused boxlist_iou + Matcher: gt [G,4], boxes [A,4] xyxy -> matches [A] int64 (gt index, -1 below, -2 between)."""
    _need_cuda(gt, boxes)
    L = _l.lib()
    gt, boxes = _f32c(gt), _f32c(boxes)
    G, A = gt.shape[0], boxes.shape[0]
    out = torch.empty(A, dtype=torch.int64, device=boxes.device)
    wsb = L.mmt_match_workspace_bytes(G, A)
    ws = torch.empty(wsb, dtype=torch.uint8, device=boxes.device)
    check(L.mmt_match_boxes_f32(_p(gt), G, _p(boxes), A, float(high_threshold), float(low_threshold),
                                int(bool(allow_low_quality_matches)), _p(out), _p(ws), wsb, _stream()), "match_boxes")
    return out


def encode_boxes(gt, matched_idx, boxes, weights):
    """targets[t] = BoxCoder(weights).encode(gt[max(matched_idx[t], 0)], boxes[t]) in one launch."""
    _need_cuda(gt, boxes, matched_idx)
    L = _l.lib()
    gt, boxes = _f32c(gt), _f32c(boxes)
    if matched_idx is not None:
        matched_idx = matched_idx.to(torch.int64).contiguous()
        assert matched_idx.shape[0] == boxes.shape[0]
    out = torch.empty_like(boxes)
    wx, wy, ww, wh = weights
    check(L.mmt_encode_boxes_f32(_p(gt), _p(matched_idx), _p(boxes), boxes.shape[0], float(wx), float(wy), float(ww),
                                 float(wh), _p(out), _stream()), "encode_boxes")
    return out


def fpn_rois(boxes_cat, image_counts, k_min, k_max, canonical_scale=224.0, canonical_level=4.0, eps=1e-6):
    """boxes_cat [K,4] (per-image lists concatenated), image_counts python ints -> (rois [K,5], levels [K] int32)."""
    _need_cuda(boxes_cat)
    L = _l.lib()
    boxes_cat = _f32c(boxes_cat)
    Kn = boxes_cat.shape[0]
    assert sum(image_counts) == Kn
    rois = torch.empty((Kn, 5), dtype=torch.float32, device=boxes_cat.device)
    levels = torch.empty(Kn, dtype=torch.int32, device=boxes_cat.device)
    counts = (c_int * len(image_counts))(*[int(c) for c in image_counts])
    check(L.mmt_fpn_rois_f32(_p(boxes_cat), counts, len(image_counts), float(k_min), float(k_max), float(canonical_scale),
                             float(canonical_level), float(eps), _p(rois), _p(levels), _stream()), "fpn_rois")
    return rois, levels


def balanced_sample(labels, offsets_dev, n_images, batch_size_per_image, quota_pos, seed):
    """labels [T] fp32 (>=1 pos, 0 neg, else ignored), offsets_dev int32 [n_images+1] -> (pos_mask, neg_mask) bool [T]."""
    _need_cuda(labels, offsets_dev)
    L = _l.lib()
    labels = _f32c(labels)
    T = labels.shape[0]
    pos = torch.empty(T, dtype=torch.bool, device=labels.device)
    neg = torch.empty(T, dtype=torch.bool, device=labels.device)
    check(L.mmt_balanced_sample_f32(_p(labels), _p(offsets_dev), n_images, int(batch_size_per_image), int(quota_pos),
                                    int(seed) & 0xFFFFFFFFFFFFFFFF, _p(pos), _p(neg), _stream()), "balanced_sample")
    return pos, neg


# ---------------------------------------------------------------------------------- ROIAlign
def roi_align_forward(input, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """`_C.roi_align_forward` (reference csrc/ROIAlign.h:11-26): NCHW in -> [K,C,ph,pw] out.
    A channels_last input is consumed as is (coalesced NHWC path) and yields a channels_last output."""
    _need_cuda(input, rois)
    L = _l.lib()
    if input.dtype != torch.float32:
        input = input.float()
    nhwc = input.is_contiguous(memory_format=torch.channels_last) and not input.is_contiguous()
    if not nhwc:
        input = input.contiguous()
    rois = _f32c(rois)
    B, C, H, W = input.shape
    K = rois.shape[0]
    out = torch.empty((K, C, pooled_h, pooled_w), dtype=torch.float32, device=input.device,
                      memory_format=torch.channels_last if nhwc else torch.contiguous_format)
    if K == 0:
        return out
    check(L.mmt_roi_align_forward_f32(_p(input), _strides(input, (0, 1, 2, 3)), B, C, H, W, _p(rois), K,
                                      float(spatial_scale), pooled_h, pooled_w, sampling_ratio, _p(out),
                                      _strides(out, (0, 1, 2, 3)), _stream()), "roi_align_forward")
    return out


def roi_align_backward(grad, rois, spatial_scale, pooled_h, pooled_w, batch, channels, height, width,
                       sampling_ratio):
    """`_C.roi_align_backward` (reference csrc/ROIAlign.h:28-45)."""
    _need_cuda(grad, rois)
    L = _l.lib()
    nhwc = grad.is_contiguous(memory_format=torch.channels_last) and not grad.is_contiguous()
    grad = grad.float() if grad.dtype != torch.float32 else grad
    if not nhwc:
        grad = grad.contiguous()
    rois = _f32c(rois)
    gin = torch.empty((batch, channels, height, width), dtype=torch.float32, device=grad.device,
                      memory_format=torch.channels_last if nhwc else torch.contiguous_format).zero_()
    K = rois.shape[0]
    if K == 0:
        return gin
    check(L.mmt_roi_align_backward_f32(_p(grad), _strides(grad, (0, 1, 2, 3)), _p(rois), K,
                                       float(spatial_scale), pooled_h, pooled_w, sampling_ratio, _p(gin),
                                       _strides(gin, (0, 1, 2, 3)), batch, channels, height, width,
                                       _stream()), "roi_align_backward")
    return gin


def _level_tables(feats, scales):
    n = len(feats)
    ptrs = (c_void_p * n)(*[f.data_ptr() for f in feats])
    hs = (c_int * n)(*[f.shape[1] for f in feats])
    ws = (c_int * n)(*[f.shape[2] for f in feats])
    sc = (c_float * n)(*[float(s) for s in scales])
    return ptrs, hs, ws, sc


def roi_align_fpn_forward(feats_nhwc, scales, rois, levels, pooled, sampling_ratio):
    """Fused FPN pooler: feats_nhwc list of dense [B,h,w,C]; rois [K,5]; levels [K] int32.
    Returns dense NHWC [K, pooled, pooled, C]."""
    _need_cuda(rois, levels, *feats_nhwc)
    L = _l.lib()
    for f in feats_nhwc:
        assert f.is_contiguous() and f.dtype == torch.float32
    rois = _f32c(rois)
    levels = levels.to(torch.int32).contiguous()
    B, _, _, C = feats_nhwc[0].shape
    K = rois.shape[0]
    out = torch.empty((K, pooled, pooled, C), dtype=torch.float32, device=rois.device)
    if K == 0:
        return out
    ptrs, hs, ws, sc = _level_tables(feats_nhwc, scales)
    check(L.mmt_roi_align_fpn_forward_f32(ptrs, hs, ws, sc, len(feats_nhwc), B, C, _p(rois), K, pooled,
                                          pooled, sampling_ratio, _p(levels), _p(out), _stream()),
          "roi_align_fpn_forward")
    return out


def roi_align_fpn_backward(grad_out, scales, rois, levels, pooled, sampling_ratio, grads_nhwc):
    """Accumulates into the (caller zero-filled or live) per-level gradient maps `grads_nhwc`."""
    _need_cuda(grad_out, rois, levels, *grads_nhwc)
    L = _l.lib()
    grad_out = _f32c(grad_out)
    rois = _f32c(rois)
    levels = levels.to(torch.int32).contiguous()
    K = rois.shape[0]
    if K == 0:
        return
    B, _, _, C = grads_nhwc[0].shape
    ptrs, hs, ws, sc = _level_tables(grads_nhwc, scales)
    check(L.mmt_roi_align_fpn_backward_f32(_p(grad_out), _p(rois), _p(levels), K, pooled, pooled,
                                           sampling_ratio, ptrs, hs, ws, sc, len(grads_nhwc), B, C,
                                           _stream()), "roi_align_fpn_backward")


def roi_pool_forward(input, rois, spatial_scale, pooled_h, pooled_w):
    """`_C.roi_pool_forward` (reference csrc/ROIPool.h:10-24): returns (output, argmax int32)."""
    _need_cuda(input, rois)
    L = _l.lib()
    input = _f32c(input)
    rois = _f32c(rois)
    B, C, H, W = input.shape
    K = rois.shape[0]
    out = torch.empty((K, C, pooled_h, pooled_w), dtype=torch.float32, device=input.device)
    arg = torch.empty((K, C, pooled_h, pooled_w), dtype=torch.int32, device=input.device)
    check(L.mmt_roi_pool_forward_f32(_p(input), B, C, H, W, _p(rois), K, float(spatial_scale), pooled_h,
                                     pooled_w, _p(out), _p(arg), _stream()), "roi_pool_forward")
    return out, arg


def roi_pool_backward(grad, input, rois, argmax, spatial_scale, pooled_h, pooled_w, batch, channels,
                      height, width):
    """`_C.roi_pool_backward` (reference csrc/ROIPool.h:26-45)."""
    _need_cuda(grad, rois, argmax)
    L = _l.lib()
    grad = _f32c(grad)
    rois = _f32c(rois)
    gin = torch.zeros((batch, channels, height, width), dtype=torch.float32, device=grad.device)
    check(L.mmt_roi_pool_backward_f32(_p(grad), _p(argmax.contiguous()), _p(rois), rois.shape[0], pooled_h,
                                      pooled_w, _p(gin), batch, channels, height, width, _stream()),
          "roi_pool_backward")
    return gin


# ------------------------------------------------------------------------------- focal loss
def sigmoid_focal_loss_forward(logits, targets, num_classes, gamma, alpha):
    _need_cuda(logits, targets)
    L = _l.lib()
    logits = _f32c(logits)
    targets = targets.to(torch.int32).contiguous()
    out = torch.empty_like(logits)
    check(L.mmt_sigmoid_focal_loss_forward_f32(_p(logits), _p(targets), logits.shape[0], num_classes,
                                               float(gamma), float(alpha), _p(out), _stream()), "focal_fwd")
    return out


def sigmoid_focal_loss_backward(logits, targets, d_losses, num_classes, gamma, alpha):
    _need_cuda(logits, targets, d_losses)
    L = _l.lib()
    logits = _f32c(logits)
    d_losses = _f32c(d_losses)
    targets = targets.to(torch.int32).contiguous()
    out = torch.empty_like(logits)
    check(L.mmt_sigmoid_focal_loss_backward_f32(_p(logits), _p(targets), _p(d_losses), logits.shape[0],
                                                num_classes, float(gamma), float(alpha), _p(out), _stream()),
          "focal_bwd")
    return out


# ---------------------------------------------------------------------------- colour views
def color_views(images_u8, params, pixel_mean):
    """images_u8 CUDA uint8 [B,H,W,3]; params[v][b] = dict(brightness, contrast, hue, rects=[(top, left, fill[h,w,3])]).
    Returns a list over views of fp32 [B,3,H,W] (BGR * 255 - mean), bit-exact with the reference's PIL pipeline."""
    import numpy as np
    _need_cuda(images_u8)
    L = _l.lib()
    assert images_u8.dtype == torch.uint8 and images_u8.is_contiguous() and images_u8.shape[-1] == 3
    B, H, W, _ = images_u8.shape
    V = len(params)
    dev = images_u8.device
    flat = [p for view in params for p in view]
    assert len(flat) == V * B
    br = torch.tensor([p["brightness"] for p in flat], dtype=torch.float32).to(dev)
    ct = torch.tensor([p["contrast"] for p in flat], dtype=torch.float32).to(dev)
    hs = torch.tensor([int(np.trunc(p["hue"] * 255.0)) % 256 for p in flat], dtype=torch.int32).to(dev)
    rec, pool, off = [], [], 0
    for vb, p in enumerate(flat):
        for top, left, fill in p["rects"]:
            h, w = fill.shape[:2]
            if h == 0 or w == 0:
                continue
            rec.append((vb, int(top), int(left), int(h), int(w), off))
            blk = np.ascontiguousarray(fill).astype(np.uint8).reshape(-1)   # numpy's float -> uint8 assignment: truncation
            pool.append(blk)
            off += blk.size
    rects_dev = pool_dev = None
    if rec:
        arr = np.zeros(len(rec), dtype=np.dtype([("vb", "<i4"), ("top", "<i4"), ("left", "<i4"), ("h", "<i4"), ("w", "<i4"),
                                                  ("pad", "<i4"), ("fill", "<i8")]))
        for i, r in enumerate(rec):
            arr[i] = (r[0], r[1], r[2], r[3], r[4], 0, r[5])
        rects_dev = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)
        pool_dev = torch.from_numpy(np.concatenate(pool)).to(dev)
    sums = torch.empty(V * B, dtype=torch.int64, device=dev)
    out = torch.empty((V, B, 3, H, W), dtype=torch.float32, device=dev)
    check(L.mmt_color_luma_sums_u8(_p(images_u8), _p(br), V, B, H, W, _p(sums), _stream()), "color_luma_sums")
    mb, mg, mr = (float(m) for m in pixel_mean)
    check(L.mmt_color_views_u8(_p(images_u8), _p(br), _p(ct), _p(hs), _p(sums), V, B, H, W, _p(rects_dev), len(rec),
                               _p(pool_dev), mb, mg, mr, _p(out), _stream()), "color_views")
    return [out[v] for v in range(V)]


# -------------------------------------------------------------------------------- RPN loss
def _rpn_level_tables(fused):
    n = len(fused)
    ptrs = (c_void_p * n)(*[f.data_ptr() for f in fused])
    pix = (ctypes.c_int64 * n)(*[f.shape[1] * f.shape[2] for f in fused])
    return ptrs, pix


def rpn_loss_forward(fused, num_anchors, labels, targets, pos, neg, beta):
    """fused: per-level head maps [N,h,w,Cp]; returns sums [2] (BCE over sampled, smooth-L1 over positives)."""
    _need_cuda(labels, targets, pos, neg, *fused)
    L = _l.lib()
    for f in fused:
        assert f.is_contiguous() and f.dtype == torch.float32
    assert labels.is_contiguous() and targets.is_contiguous() and pos.is_contiguous() and neg.is_contiguous()
    assert labels.dtype == torch.float32 and targets.dtype == torch.float32 and pos.element_size() == 1 and neg.element_size() == 1
    N, Cp = fused[0].shape[0], fused[0].shape[3]
    assert labels.numel() == N * sum(f.shape[1] * f.shape[2] * num_anchors for f in fused)
    sums = torch.empty(2, dtype=torch.float32, device=labels.device)
    ptrs, pix = _rpn_level_tables(fused)
    check(L.mmt_rpn_loss_forward_f32(ptrs, pix, len(fused), num_anchors, Cp, N, _p(labels), _p(targets), _p(pos),
                                     _p(neg), float(beta), _p(sums), _stream()), "rpn_loss_forward")
    return sums


def rpn_loss_backward(fused, num_anchors, labels, targets, pos, neg, beta, n_sampled, g_obj, g_box):
    L = _l.lib()
    N, Cp = fused[0].shape[0], fused[0].shape[3]
    dfused = [torch.empty_like(f) for f in fused]
    ptrs, pix = _rpn_level_tables(fused)
    dptrs = (c_void_p * len(fused))(*[d.data_ptr() for d in dfused])
    check(L.mmt_rpn_loss_backward_f32(ptrs, dptrs, pix, len(fused), num_anchors, Cp, N, _p(labels), _p(targets), _p(pos),
                                      _p(neg), float(beta), _p(n_sampled), _p(g_obj), _p(g_box), _stream()),
          "rpn_loss_backward")
    return dfused


# ------------------------------------------------------------------------------------- PSM
def psm_loss(student_logits, teacher_logits, labels, temperature=0.5, sharpen=True, hard_negative=True,
             neg_weight=1.5, neg_order=None):
    """student [n,C]; teacher [V,n,C]; labels [n] int64. Returns (loss[1], grad_student[n,C], selected[n] int32)."""
    _need_cuda(student_logits, teacher_logits, labels)
    L = _l.lib()
    s = _f32c(student_logits)
    t = _f32c(teacher_logits)
    labels = labels.to(torch.int64).contiguous()
    V, n, C = t.shape
    loss = torch.empty(1, dtype=torch.float32, device=s.device)
    grad = torch.empty_like(s)
    sel = torch.zeros(max(n, 1), dtype=torch.int32, device=s.device)
    wsb = L.mmt_psm_workspace_bytes(n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=s.device)
    if neg_order is not None:
        neg_order = neg_order.to(torch.int32).contiguous()
    check(L.mmt_psm_loss_f32(_p(s), _p(t), V, _p(labels), n, C, float(temperature), int(bool(sharpen)),
                             int(bool(hard_negative)), float(neg_weight), _p(neg_order), _p(loss), _p(grad),
                             _p(sel), _p(ws), wsb, _stream()), "psm_loss")
    return loss, grad, sel[:n]


# ------------------------------------------------------------------------------------- MGD
def mgd_masks(integral_mask, level_sizes):
    """integral_mask [B,H,W] float; level_sizes list of (h,w). Returns (list of uint8 [B,h,w], sums [n_levels])."""
    _need_cuda(integral_mask)
    L = _l.lib()
    m = _f32c(integral_mask)
    B, H, W = m.shape
    n = len(level_sizes)
    total = sum(B * h * w for h, w in level_sizes)
    masks = torch.empty(total, dtype=torch.uint8, device=m.device)
    sums = torch.empty(n, dtype=torch.float32, device=m.device)
    hs = (c_int * n)(*[h for h, _ in level_sizes])
    ws = (c_int * n)(*[w for _, w in level_sizes])
    check(L.mmt_mgd_masks_f32(_p(m), B, H, W, hs, ws, n, _p(masks), _p(sums), _stream()), "mgd_masks")
    out, off = [], 0
    for h, w in level_sizes:
        out.append(masks[off: off + B * h * w].view(B, h, w))
        off += B * h * w
    return out, sums


def mgd_level(student_nhwc, teacher_views_nhwc, flip_mask, level_mask, level_mask_sum, term_weight, partial,
              want_grad):
    """One pyramid level. student / teachers dense NHWC [B,h,w,C]; `partial` [n_views] accumulated."""
    _need_cuda(student_nhwc, level_mask, level_mask_sum, partial, *teacher_views_nhwc)
    L = _l.lib()
    B, H, W, C = student_nhwc.shape
    V = len(teacher_views_nhwc)
    for t in teacher_views_nhwc:
        assert t.is_contiguous() and t.shape == student_nhwc.shape and t.dtype == torch.float32
    assert student_nhwc.is_contiguous() and student_nhwc.dtype == torch.float32
    st = Strides4(H * W * C, 1, W * C, C)
    ptrs = (c_void_p * V)(*[t.data_ptr() for t in teacher_views_nhwc])
    grad = torch.empty_like(student_nhwc) if want_grad else None
    check(L.mmt_mgd_level_f32(_p(student_nhwc), st, ptrs, st, V, int(flip_mask), _p(level_mask),
                              _p(level_mask_sum), B, C, H, W, float(term_weight), _p(partial), _p(grad), st,
                              _stream()), "mgd_level")
    return grad


def mgd_finalize(partials, mask_sums, n_levels, n_views, channels):
    _need_cuda(partials, mask_sums)
    L = _l.lib()
    loss = torch.empty(1, dtype=torch.float32, device=partials.device)
    check(L.mmt_mgd_finalize_f32(_p(partials), _p(mask_sums), n_levels, n_views, channels, _p(loss), _stream()),
          "mgd_finalize")
    return loss


def rasterize_polygons(packed, roi_boxes, roi_instance, mask_size):
    """packed = SegmentationMask.packed(device); roi_boxes [K,4]; roi_instance [K] -> [K,M,M] fp32 {0,1}."""
    xy, poly_off, inst_off = packed
    _need_cuda(xy, poly_off, inst_off, roi_boxes, roi_instance)
    L = _l.lib()
    roi_boxes = _f32c(roi_boxes)
    roi_instance = roi_instance.to(torch.int32).contiguous()
    K = roi_boxes.shape[0]
    out = torch.empty((K, mask_size, mask_size), dtype=torch.float32, device=roi_boxes.device)
    check(L.mmt_rasterize_polygons_f32(_p(xy), _p(poly_off), _p(inst_off), _p(roi_boxes), _p(roi_instance), K,
                                       mask_size, _p(out), _stream()), "rasterize_polygons")
    return out


def paste_masks_sum(mask_probs, boxes, image_index, n_images, im_h, im_w, threshold=0.5):
    """mask_probs [D,M,M]; boxes [D,4]; image_index [D] int32 -> integral foreground map [n_images, im_h, im_w]."""
    _need_cuda(mask_probs, boxes, image_index)
    L = _l.lib()
    mask_probs = _f32c(mask_probs)
    boxes = _f32c(boxes)
    image_index = image_index.to(torch.int32).contiguous()
    out = torch.empty((n_images, im_h, im_w), dtype=torch.float32, device=boxes.device)
    check(L.mmt_paste_masks_sum_f32(_p(mask_probs), _p(boxes), _p(image_index), mask_probs.shape[0],
                                    mask_probs.shape[-1], n_images, im_h, im_w, float(threshold), _p(out), _stream()),
          "paste_masks_sum")
    return out


# ------------------------------------------------------------------------------- EMA / SGD
def ema_(teacher_flat, student_flat, alpha):
    _need_cuda(teacher_flat, student_flat)
    L = _l.lib()
    assert teacher_flat.is_contiguous() and student_flat.is_contiguous()
    assert teacher_flat.dtype == torch.float32 and student_flat.dtype == torch.float32
    assert teacher_flat.numel() == student_flat.numel()
    check(L.mmt_ema_f32(_p(teacher_flat), _p(student_flat), teacher_flat.numel(), float(alpha), _stream()), "ema")


def sgd_ema_(student_flat, teacher_flat, grad_flat, momentum_flat, segments_dev, n_segments, momentum, alpha,
             first_step, grad_scale=1.0):
    _need_cuda(student_flat, grad_flat, momentum_flat, segments_dev)
    L = _l.lib()
    check(L.mmt_sgd_ema_f32(_p(student_flat), _p(teacher_flat), _p(grad_flat), _p(momentum_flat),
                            student_flat.numel(), _p(segments_dev), n_segments, float(momentum), float(alpha),
                            int(bool(first_step)), float(grad_scale), _stream()), "sgd_ema")


def make_segments(entries, device):
    """entries: list of (begin, end, lr, weight_decay, trainable). Returns a device uint8 tensor."""
    arr = (_l.ParamSegment * len(entries))()
    for i, (b, e, lr, wd, tr) in enumerate(entries):
        arr[i].begin, arr[i].end, arr[i].lr, arr[i].weight_decay, arr[i].trainable = b, e, lr, wd, int(tr)
    raw = bytes(arr)
    t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
    if torch.device(device).type == "cuda":    # pinned staging: a pageable H2D copy would synchronise the stream (the table is
        t = t.pin_memory()                     # rebuilt whenever the learning rate changes: every iteration during warm-up)
    return t.to(device, non_blocking=True)


# ------------------------------------------------------------------------------------ convs
def conv_shape(batch, in_h, in_w, in_c, out_c, k, stride, pad):
    oh = (in_h + 2 * pad - k) // stride + 1
    ow = (in_w + 2 * pad - k) // stride + 1
    return ConvShape(batch, in_h, in_w, in_c, out_c, k, k, stride, pad, oh, ow)


def conv_force_cta_group(cta_group):
    """0 = per-problem cost model (default); 1 = one CTA per 128 x BN tile; 2 = CTA pairs (tcgen05 cta_group::2)."""
    check(_l.lib().mmt_conv_force_cta_group(int(cta_group)), "conv_force_cta_group")


def conv_fprop(x, w, shape, scale=None, bias=None, residual=None, relu=False, impl=None, out=None):
    """x NHWC [B,H,W,Ci] dense, w OHWI [Co,kh,kw,Ci] dense -> y NHWC [B,OH,OW,Co]."""
    _need_cuda(x, w, scale, bias, residual)
    L = _l.lib()
    impl = _conv_impl if impl is None else impl
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.float32 and w.dtype == torch.float32
    y = out if out is not None else torch.empty((shape.batch, shape.out_h, shape.out_w, shape.out_c),
                                                dtype=torch.float32, device=x.device)
    check(L.mmt_conv_fprop_f32(_p(x), _p(w), _p(scale), _p(bias), _p(residual), _p(y), ctypes.byref(shape),
                               EPI_RELU if relu else 0, impl, _stream()), "conv_fprop")
    return y


STEM_INPUT_NCHW = 2


def conv_stem_fprop(x, w, shape, scale=None, bias=None, relu=False, nchw=False):
    """7x7 / stride-2 / pad-3 stem over 3 channels on the tensor cores (space-to-depth implicit GEMM).
    x NHWC [B,H,W,3] (or NCHW [B,3,H,W] with nchw=True), w OHWI [Co,7,7,3] -> y NHWC [B,ceil(H/2),ceil(W/2),Co]."""
    _need_cuda(x, w, scale, bias)
    L = _l.lib()
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.float32 and w.dtype == torch.float32
    wsb = L.mmt_conv_stem_workspace_bytes(ctypes.byref(shape))
    if wsb == 0:
        raise RuntimeError("conv_stem_fprop: unsupported stem shape")
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    y = torch.empty((shape.batch, shape.out_h, shape.out_w, shape.out_c), dtype=torch.float32, device=x.device)
    check(L.mmt_conv_stem_fprop_f32(_p(x), _p(w), _p(scale), _p(bias), _p(y), ctypes.byref(shape),
                                    (EPI_RELU if relu else 0) | (STEM_INPUT_NCHW if nchw else 0), _p(ws), wsb, _stream()),
          "conv_stem_fprop")
    return y


def conv_dgrad(dy, w, shape, impl=None):
    _need_cuda(dy, w)
    L = _l.lib()
    impl = _conv_impl if impl is None else impl
    assert dy.is_contiguous() and w.is_contiguous()
    dx = torch.empty((shape.batch, shape.in_h, shape.in_w, shape.in_c), dtype=torch.float32, device=dy.device)
    wsb = L.mmt_conv_workspace_bytes(ctypes.byref(shape))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dy.device)
    check(L.mmt_conv_dgrad_f32(_p(dy), _p(w), _p(dx), ctypes.byref(shape), impl, _p(ws), wsb, _stream()),
          "conv_dgrad")
    return dx


def conv_wgrad(x, dy, shape, want_dbias=False, impl=None):
    _need_cuda(x, dy)
    L = _l.lib()
    impl = _conv_impl if impl is None else impl
    assert x.is_contiguous() and dy.is_contiguous()
    dw = torch.empty((shape.out_c, shape.k_h, shape.k_w, shape.in_c), dtype=torch.float32, device=x.device)
    db = torch.empty(shape.out_c, dtype=torch.float32, device=x.device) if want_dbias else None
    wsb = L.mmt_conv_workspace_bytes(ctypes.byref(shape))
    ws = torch.empty(wsb, dtype=torch.uint8, device=x.device)
    check(L.mmt_conv_wgrad_f32(_p(x), _p(dy), _p(dw), _p(db), ctypes.byref(shape), impl, _p(ws), wsb, _stream()),
          "conv_wgrad")
    return dw, db


def conv_wgrad_accumulate_(x, dy, shape, dw):
    """dw += wgrad(x, dy) in place (tcgen05 path).  Returns False — nothing done — when the shape or the conv math mode
    needs the generic path."""
    if _conv_impl != CONV_AUTO or not (dw.is_contiguous() and dw.dtype == torch.float32):
        return False
    _need_cuda(x, dy, dw)
    L = _l.lib()
    assert x.is_contiguous() and dy.is_contiguous()
    rc = L.mmt_conv_wgrad_accumulate_f32(_p(x), _p(dy), _p(dw), ctypes.byref(shape), _stream())
    if rc == 4:  # MMT_ERR_UNSUPPORTED
        return False
    check(rc, "conv_wgrad_accumulate")
    return True


DGRAD_AUX_NONE, DGRAD_AUX_ADD, DGRAD_AUX_MASK = 0, 1, 2


def conv_dgrad_fused(dy, w, shape, k_scale=None, aux=None, aux_mode=DGRAD_AUX_NONE):
    """dx = E(conv_T(dy * k_scale[out_c], w)) with E = identity | + aux | masked by (aux > 0): the data gradient of a
    Conv -> FrozenBN chain given the gradient of the BN output, with the skip-connection add or the previous layer's ReLU
    backward in the epilogue (mmt_conv_dgrad_fused_f32).  In fp32 conv math (parity mode) or for shapes outside the tcgen05
    kernel the same result is composed from the unfused pieces."""
    _need_cuda(dy, w, k_scale, aux)
    L = _l.lib()
    assert dy.is_contiguous() and w.is_contiguous() and (aux is None or aux.is_contiguous())
    if _conv_impl == CONV_AUTO:
        dx = torch.empty((shape.batch, shape.in_h, shape.in_w, shape.in_c), dtype=torch.float32, device=dy.device)
        wsb = L.mmt_conv_workspace_bytes(ctypes.byref(shape))
        ws = torch.empty(wsb, dtype=torch.uint8, device=dy.device)
        rc = L.mmt_conv_dgrad_fused_f32(_p(dy), _p(w), _p(k_scale), _p(aux), int(aux_mode), _p(dx), ctypes.byref(shape),
                                        _p(ws), wsb, _stream())
        if rc != 4:  # MMT_ERR_UNSUPPORTED -> composition below
            check(rc, "conv_dgrad_fused")
            return dx
    g = dy if k_scale is None else dy * k_scale
    dx = conv_dgrad(g, w, shape)
    if aux_mode == DGRAD_AUX_ADD:
        dx += aux
    elif aux_mode == DGRAD_AUX_MASK:
        dx = dx * (aux > 0)
    return dx


def conv_wgrad_scaled(x, dy, shape, row_scale=None, into=None):
    """row_scale[o] * wgrad(x, dy)[o]: accumulated into `into` (a live gradient buffer; returns None) or returned."""
    _need_cuda(x, dy, row_scale, into)
    L = _l.lib()
    assert x.is_contiguous() and dy.is_contiguous()
    if _conv_impl == CONV_AUTO and (into is None or (into.is_contiguous() and into.dtype == torch.float32)):
        dw = into if into is not None else torch.empty((shape.out_c, shape.k_h, shape.k_w, shape.in_c),
                                                       dtype=torch.float32, device=x.device)
        rc = L.mmt_conv_wgrad_scaled_f32(_p(x), _p(dy), _p(row_scale), _p(dw), ctypes.byref(shape),
                                         1 if into is not None else 0, _stream())
        if rc != 4:
            check(rc, "conv_wgrad_scaled")
            return None if into is not None else dw
    g = dy if row_scale is None else dy * row_scale
    dw, _ = conv_wgrad(x, g, shape)
    if into is not None:
        into += dw
        return None
    return dw


def bn_relu_backward(grad_y, y, scale, relu):
    _need_cuda(grad_y, y, scale)
    L = _l.lib()
    assert grad_y.is_contiguous()
    C = grad_y.shape[-1]
    gz = torch.empty_like(grad_y)
    check(L.mmt_bn_relu_backward_f32(_p(grad_y), _p(y), _p(scale), _p(gz), grad_y.numel() // C, C,
                                     int(bool(relu)), _stream()), "bn_relu_backward")
    return gz


def maxpool3x3s2(x):
    _need_cuda(x)
    L = _l.lib()
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=torch.float32, device=x.device)
    check(L.mmt_maxpool3x3s2_forward_f32(_p(x), _p(y), B, H, W, C, _stream()), "maxpool")
    return y


def subsample2(x):
    _need_cuda(x)
    L = _l.lib()
    B, H, W, C = x.shape
    y = torch.empty((B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, C), dtype=torch.float32, device=x.device)
    check(L.mmt_subsample2_forward_f32(_p(x), _p(y), B, H, W, C, _stream()), "subsample2")
    return y


def subsample2_backward_(grad_y, grad_x):
    _need_cuda(grad_y, grad_x)
    L = _l.lib()
    B, H, W, C = grad_x.shape
    check(L.mmt_subsample2_backward_f32(_p(grad_y), _p(grad_x), B, H, W, C, _stream()), "subsample2_backward")


def upsample2_add(lateral, top):
    _need_cuda(lateral, top)
    L = _l.lib()
    B, H, W, C = lateral.shape
    out = torch.empty_like(lateral)
    check(L.mmt_upsample2_add_f32(_p(lateral), _p(top), _p(out), B, H, W, top.shape[1], top.shape[2], C, _stream()),
          "upsample2_add")
    return out


def upsample2_backward_(grad_out, grad_top):
    _need_cuda(grad_out, grad_top)
    L = _l.lib()
    B, H, W, C = grad_out.shape
    check(L.mmt_upsample2_backward_f32(_p(grad_out), _p(grad_top), B, H, W, grad_top.shape[1], grad_top.shape[2],
                                       C, _stream()), "upsample2_backward")


def nchw_to_nhwc(x):
    _need_cuda(x)
    L = _l.lib()
    x = _f32c(x)
    N, C, H, W = x.shape
    y = torch.empty((N, H, W, C), dtype=torch.float32, device=x.device)
    check(L.mmt_nchw_to_nhwc_f32(_p(x), _p(y), N, C, H, W, _stream()), "nchw_to_nhwc")
    return y


def nhwc_to_nchw(x):
    _need_cuda(x)
    L = _l.lib()
    x = _f32c(x)
    N, H, W, C = x.shape
    y = torch.empty((N, C, H, W), dtype=torch.float32, device=x.device)
    check(L.mmt_nhwc_to_nchw_f32(_p(x), _p(y), N, C, H, W, _stream()), "nhwc_to_nchw")
    return y
