"""ctypes binding of libmmtpsm_b200.so (the C ABI declared in include/mmtpsm_b200.h).

This is the ONLY place the shared library is opened.  There is no CPU or PyTorch fallback: if the
library is missing, `lib()` raises, and every op in `kernels.py` refuses non-CUDA tensors.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int64, c_size_t, c_uint, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_lib", "libmmtpsm_b200.so")


class Strides4(Structure):
    _fields_ = [("n", c_int64), ("c", c_int64), ("h", c_int64), ("w", c_int64)]


class ConvShape(Structure):
    _fields_ = [("batch", c_int), ("in_h", c_int), ("in_w", c_int), ("in_c", c_int),
                ("out_c", c_int), ("k_h", c_int), ("k_w", c_int), ("stride", c_int), ("pad", c_int),
                ("out_h", c_int), ("out_w", c_int)]


class ParamSegment(Structure):
    _fields_ = [("begin", c_int64), ("end", c_int64), ("lr", c_float), ("weight_decay", c_float),
                ("trainable", c_int), ("reserved", c_int)]


P = c_void_p
_PROTOS = {
    "mmt_last_error": (c_char_p, []),
    "mmt_version": (c_char_p, []),
    "mmt_launch_count": (c_int64, []),
    "mmt_reset_launch_count": (None, []),
    "mmt_add_launch_count": (None, [c_int64]),
    "mmt_prof_enable": (None, [c_int]),
    "mmt_prof_collect": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(ctypes.c_double),
                                 POINTER(c_int64), c_int]),
    "mmt_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mmt_nms_segmented_f32": (c_int, [P, P, POINTER(c_int), c_int, c_float, P, P, P, c_size_t, P]),
    "mmt_nms_segmented_flags_f32": (c_int, [P, P, POINTER(c_int), c_int, c_float, P, P, P, P, c_size_t, P]),
    "mmt_rpn_gather_f32": (c_int, [P, c_int, c_int, c_int64, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "mmt_rpn_topk_gather_f32": (c_int, [POINTER(P), POINTER(P), POINTER(P), POINTER(c_int), POINTER(c_int), c_int, c_int, c_int,
                                        c_int, c_int, P, P, P, P]),
    "mmt_nms_f32": (c_int, [P, P, c_int, c_float, P, P, P, c_size_t, P]),
    "mmt_decode_clip_boxes_f32": (c_int, [P, P, P, P, c_int, c_int, c_float, c_float, c_float, c_float, c_float,
                                          c_float, P, P, P, P]),
    "mmt_match_workspace_bytes": (c_size_t, [c_int, c_int]),
    "mmt_match_boxes_f32": (c_int, [P, c_int, P, c_int, c_float, c_float, c_int, P, P, c_size_t, P]),
    "mmt_encode_boxes_f32": (c_int, [P, P, P, c_int64, c_float, c_float, c_float, c_float, P, P]),
    "mmt_fpn_rois_f32": (c_int, [P, POINTER(c_int), c_int, c_float, c_float, c_float, c_float, c_float, P, P, P]),
    "mmt_balanced_sample_f32": (c_int, [P, P, c_int, c_int, c_int, ctypes.c_uint64, P, P, P]),
    "mmt_roi_align_forward_f32": (c_int, [P, Strides4, c_int, c_int, c_int, c_int, P, c_int, c_float,
                                          c_int, c_int, c_int, P, Strides4, P]),
    "mmt_roi_align_backward_f32": (c_int, [P, Strides4, P, c_int, c_float, c_int, c_int, c_int, P,
                                           Strides4, c_int, c_int, c_int, c_int, P]),
    "mmt_roi_align_fpn_forward_f32": (c_int, [POINTER(P), POINTER(c_int), POINTER(c_int),
                                              POINTER(c_float), c_int, c_int, c_int, P, c_int, c_int,
                                              c_int, c_int, P, P, P]),
    "mmt_roi_align_fpn_backward_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, POINTER(P),
                                               POINTER(c_int), POINTER(c_int), POINTER(c_float), c_int,
                                               c_int, c_int, P]),
    "mmt_roi_pool_forward_f32": (c_int, [P, c_int, c_int, c_int, c_int, P, c_int, c_float, c_int,
                                         c_int, P, P, P]),
    "mmt_roi_pool_backward_f32": (c_int, [P, P, P, c_int, c_int, c_int, P, c_int, c_int, c_int,
                                          c_int, P]),
    "mmt_color_luma_sums_u8": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P]),
    "mmt_color_views_u8": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, c_int, P, c_float, c_float, c_float, P, P]),
    "mmt_rpn_loss_forward_f32": (c_int, [POINTER(P), POINTER(c_int64), c_int, c_int, c_int, c_int, P, P, P, P, c_float, P, P]),
    "mmt_rpn_loss_backward_f32": (c_int, [POINTER(P), POINTER(P), POINTER(c_int64), c_int, c_int, c_int, c_int, P, P, P, P, c_float,
                                          P, P, P, P]),
    "mmt_sigmoid_focal_loss_forward_f32": (c_int, [P, P, c_int, c_int, c_float, c_float, P, P]),
    "mmt_sigmoid_focal_loss_backward_f32": (c_int, [P, P, P, c_int, c_int, c_float, c_float, P, P]),
    "mmt_psm_workspace_bytes": (c_size_t, [c_int]),
    "mmt_psm_loss_f32": (c_int, [P, P, c_int, P, c_int, c_int, c_float, c_int, c_int, c_float, P, P,
                                 P, P, P, c_size_t, P]),
    "mmt_mgd_masks_f32": (c_int, [P, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int), c_int, P,
                                  P, P]),
    "mmt_mgd_level_f32": (c_int, [P, Strides4, POINTER(P), Strides4, c_int, c_uint, P, P, c_int,
                                  c_int, c_int, c_int, c_float, P, P, Strides4, P]),
    "mmt_mgd_finalize_f32": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "mmt_rasterize_polygons_f32": (c_int, [P, P, P, P, P, c_int, c_int, P, P]),
    "mmt_paste_masks_sum_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_float, P, P]),
    "mmt_ema_f32": (c_int, [P, P, c_int64, c_float, P]),
    "mmt_sgd_ema_f32": (c_int, [P, P, P, P, c_int64, P, c_int, c_float, c_float, c_int, c_float, P]),
    "mmt_conv_set_max_ctas": (c_int, [c_int]),
    "mmt_conv_force_cta_group": (c_int, [c_int]),
    "mmt_conv_set_pdl": (c_int, [c_int]),
    "mmt_conv_set_pipeline": (c_int, [c_int, c_int]),
    "mmt_conv_dgrad_fused_f32": (c_int, [P, P, P, P, c_int, P, POINTER(ConvShape), P, c_size_t, P]),
    "mmt_conv_wgrad_scaled_f32": (c_int, [P, P, P, P, POINTER(ConvShape), c_int, P]),
    "mmt_conv_workspace_bytes": (c_size_t, [POINTER(ConvShape)]),
    "mmt_conv_fprop_f32": (c_int, [P, P, P, P, P, P, POINTER(ConvShape), c_int, c_int, P]),
    "mmt_conv_stem_workspace_bytes": (c_size_t, [POINTER(ConvShape)]),
    "mmt_conv_stem_fprop_f32": (c_int, [P, P, P, P, P, POINTER(ConvShape), c_int, P, c_size_t, P]),
    "mmt_conv_dgrad_f32": (c_int, [P, P, P, POINTER(ConvShape), c_int, P, c_size_t, P]),
    "mmt_conv_wgrad_f32": (c_int, [P, P, P, P, POINTER(ConvShape), c_int, P, c_size_t, P]),
    "mmt_conv_wgrad_accumulate_f32": (c_int, [P, P, P, POINTER(ConvShape), P]),
    "mmt_bn_relu_backward_f32": (c_int, [P, P, P, P, c_int64, c_int, c_int, P]),
    "mmt_maxpool3x3s2_forward_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "mmt_subsample2_forward_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "mmt_subsample2_backward_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "mmt_upsample2_add_f32": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mmt_upsample2_backward_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "mmt_nchw_to_nhwc_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
    "mmt_nhwc_to_nchw_f32": (c_int, [P, P, c_int, c_int, c_int, c_int, P]),
}

_lib = None


def exported_symbols():
    """Names the header declares (used by the CPU-side ABI test)."""
    return sorted(_PROTOS)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libmmtpsm_b200.so is not built (%s). Run `python mmt-psm_b200/build.py`; "
                "there is no CPU / PyTorch fallback for the hot path." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOS.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
        # experiment knobs of the conv engine (benchmarks / bisecting): MMT_CONV_PDL=0, MMT_CONV_CTA_GROUP=1|2
        if os.environ.get("MMT_CONV_PDL") in ("0", "1"):
            handle.mmt_conv_set_pdl(int(os.environ["MMT_CONV_PDL"]))
        if os.environ.get("MMT_CONV_PREFETCH"):
            handle.mmt_conv_set_pipeline(0, int(os.environ["MMT_CONV_PREFETCH"]))
        if os.environ.get("MMT_CONV_CTA_GROUP") in ("0", "1", "2"):
            handle.mmt_conv_force_cta_group(int(os.environ["MMT_CONV_CTA_GROUP"]))
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().mmt_last_error().decode("utf-8", "replace")
        raise RuntimeError("mmtpsm_b200 %s failed (status %d): %s" % (what, rc, msg))


def launch_count():
    return int(lib().mmt_launch_count())


def reset_launch_count():
    lib().mmt_reset_launch_count()


def add_launch_count(n):
    lib().mmt_add_launch_count(int(n))


PROF_FAMILIES = ("conv_fprop_tc", "conv_dgrad_tc", "conv_wgrad_tc", "conv_simt", "roi_align", "sgd_ema", "other")


def prof_enable(on=True):
    lib().mmt_prof_enable(1 if on else 0)


def prof_collect():
    """{family: dict(ms, flops, bytes, launches)} for everything recorded since prof_enable(True)."""
    n = len(PROF_FAMILIES)
    ms, fl, by = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_double * n)()
    cnt = (c_int64 * n)()
    check(lib().mmt_prof_collect(ms, fl, by, cnt, n), "prof_collect")
    return {k: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(cnt[i])) for i, k in enumerate(PROF_FAMILIES)}
