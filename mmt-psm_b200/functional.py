"""Autograd bindings of the CUDA kernels (NHWC fp32 activations, OHWI filters).

Each `torch.autograd.Function` below is one fused forward launch and, in backward, the matching dgrad /
wgrad / scatter launches of libmmtpsm_b200 — there is no PyTorch fallback inside any of them.
Layout conventions:  activations [B,H,W,C] dense; filters [Co,kh,kw,Ci] dense; ROI features [K,P,P,C].
"""
import torch

from . import kernels as K


class _ConvBnAct(torch.autograd.Function):
    """y = act(conv(x, w) * scale + bias + residual).

    Replaces Conv2d -> FrozenBatchNorm2d -> (+identity) -> relu_ (reference layers/misc.py:30-43,
    layers/batch_norm.py:19-24, modeling/backbone/resnet.py:254-275) and Conv2d(+bias)(+relu) of the FPN /
    RPN / mask heads.  `scale` is a frozen per-channel multiplier (no gradient); `bias` may be a learnable
    conv bias (gradient = per-channel sum) or a frozen BN shift (requires_grad False)."""

    @staticmethod
    def forward(ctx, x, w, scale, bias, residual, relu, stride, pad):
        B, H, W, Ci = x.shape
        Co, k = w.shape[0], w.shape[1]
        shape = K.conv_shape(B, H, W, Ci, Co, k, stride, pad)
        y = K.conv_fprop(x, w, shape, scale, bias, residual, relu)
        ctx.shape, ctx.relu = shape, relu
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, w, scale, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, scale, y = ctx.saved_tensors
        shape = ctx.shape
        gy = gy.contiguous()
        if ctx.relu and scale is not None and not ctx.has_res:
            g_res = None
            g_conv = K.bn_relu_backward(gy, y, scale, True)                    # ReLU mask and BN scale in one pass
        else:
            g_res = K.bn_relu_backward(gy, y, None, True) if ctx.relu else gy  # d/d(pre-activation)
            g_conv = K.bn_relu_backward(g_res, None, scale, False) if scale is not None else g_res
        need_x, need_w, _, need_b, need_r = ctx.needs_input_grad[:5]
        dx = K.conv_dgrad(g_conv, w, shape) if need_x else None
        dw = db = None
        if (need_w and not need_b and _fused_grad_sink is not None and w.is_leaf and w.grad is not None
                and K.conv_wgrad_accumulate_(x, g_conv, shape, w.grad)):
            _fused_grad_sink([w])     # dW went straight into the optimizer's gradient arena
            need_w = False
        elif need_w or need_b:
            dw, db = K.conv_wgrad(x, g_conv, shape, want_dbias=need_b)
        return dx, (dw if need_w else None), None, db, (g_res if (need_r and ctx.has_res) else None), None, None, None


# When set (solver.FusedSGDEMA does it), weight gradients of the tcgen05 convs are accumulated IN PLACE into `w.grad`
# (a persistent view of the flat gradient arena) and the callback is told which parameters were written — autograd
# then sees no gradient for w, so no AccumulateGrad add and no post-accumulate hook runs for it.
_fused_grad_sink = None


def set_fused_grad_sink(callback):
    global _fused_grad_sink
    _fused_grad_sink = callback


def conv_bn_act(x, w, scale=None, bias=None, residual=None, relu=False, stride=1, pad=0):
    return _ConvBnAct.apply(x, w, scale, bias, residual, relu, stride, pad)


class _Bottleneck(torch.autograd.Function):
    """One ResNet bottleneck (reference modeling/backbone/resnet.py:254-275: conv1-bn1-relu, conv2-bn2-relu, conv3-bn3,
    + identity / downsample-bn, relu) as ONE autograd node.

    Forward: the same 3 (4) fused conv launches as the layer-by-layer path.  Backward: what autograd would spread over
    ~5 elementwise passes per block (ReLU masks, frozen-BN scales, the skip-connection add) is folded into the conv
    kernels — BN scale into the transformed dgrad filter and the wgrad epilogue, the ReLU mask of each inner activation
    into the epilogue of the dgrad that produces its gradient, the skip add into conv1's dgrad epilogue
    (mmt_conv_dgrad_fused_f32 / mmt_conv_wgrad_scaled_f32); one elementwise pass (the block output's ReLU mask) remains."""

    @staticmethod
    def forward(ctx, x, w1, w2, w3, wd, s1, b1, s2, b2, s3, b3, sd, bd, stride):
        B, H, W, Cin = x.shape
        Cmid, Cout = w1.shape[0], w3.shape[0]
        sh1 = K.conv_shape(B, H, W, Cin, Cmid, 1, stride, 0)
        o1 = K.conv_fprop(x, w1, sh1, s1, b1, None, True)
        sh2 = K.conv_shape(B, sh1.out_h, sh1.out_w, Cmid, Cmid, 3, 1, 1)
        o2 = K.conv_fprop(o1, w2, sh2, s2, b2, None, True)
        shd = None
        if wd is not None:
            shd = K.conv_shape(B, H, W, Cin, Cout, 1, stride, 0)
            idt = K.conv_fprop(x, wd, shd, sd, bd, None, False)
        else:
            idt = x
        sh3 = K.conv_shape(B, sh2.out_h, sh2.out_w, Cmid, Cout, 1, 1, 0)
        y = K.conv_fprop(o2, w3, sh3, s3, b3, idt, True)
        ctx.shapes = (sh1, sh2, sh3, shd)
        ctx.save_for_backward(x, o1, o2, y, w1, w2, w3, wd, s1, s2, s3, sd)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, o1, o2, y, w1, w2, w3, wd, s1, s2, s3, sd = ctx.saved_tensors
        sh1, sh2, sh3, shd = ctx.shapes
        need = ctx.needs_input_grad
        g3 = K.bn_relu_backward(gy.contiguous(), y, None, True)          # gradient of (bn3 output + identity)
        grads = {}

        def wgrad(idx, xin, g, shape, w, scale):
            if not need[idx]:
                return
            if _fused_grad_sink is not None and w.is_leaf and w.grad is not None and w.grad.is_contiguous():
                K.conv_wgrad_scaled(xin, g, shape, scale, into=w.grad)   # straight into the optimizer's gradient arena
                _fused_grad_sink([w])
            else:
                grads[idx] = K.conv_wgrad_scaled(xin, g, shape, scale)

        d_o2 = K.conv_dgrad_fused(g3, w3, sh3, s3, o2, K.DGRAD_AUX_MASK)  # through conv3, bn3 and relu(bn2)
        wgrad(3, o2, g3, sh3, w3, s3)
        d_o1 = K.conv_dgrad_fused(d_o2, w2, sh2, s2, o1, K.DGRAD_AUX_MASK)
        wgrad(2, o1, d_o2, sh2, w2, s2)
        wgrad(1, x, d_o1, sh1, w1, s1)
        dx = None
        if wd is not None:
            wgrad(4, x, g3, shd, wd, sd)
        if need[0]:
            if wd is None:
                dx = K.conv_dgrad_fused(d_o1, w1, sh1, s1, g3, K.DGRAD_AUX_ADD)   # + the identity branch's gradient
            else:
                dxd = K.conv_dgrad_fused(g3, wd, shd, sd)
                if sh1.stride == 1:
                    dx = K.conv_dgrad_fused(d_o1, w1, sh1, s1, dxd, K.DGRAD_AUX_ADD)
                else:
                    dx = K.conv_dgrad_fused(d_o1, w1, sh1, s1)
                    dx += dxd
        return (dx, grads.get(1), grads.get(2), grads.get(3), grads.get(4)) + (None,) * 9


def bottleneck(x, w1, w2, w3, wd, s1, b1, s2, b2, s3, b3, sd, bd, stride):
    return _Bottleneck.apply(x, w1, w2, w3, wd, s1, b1, s2, b2, s3, b3, sd, bd, stride)


def linear(x, w, bias=None, relu=False):
    """x [N,K] @ w[O,K]^T (+bias)(+relu) on the conv engine: a 1x1 convolution over N 'pixels'.
    Replaces nn.Linear of the box head (reference roi_box_feature_extractors.py:97-98,119-121;
    roi_box_predictors.py:39-40)."""
    N, Kd = x.shape
    if N == 0:
        return x.new_zeros((0, w.shape[0]))
    y = conv_bn_act(x.reshape(1, 1, N, Kd), w.reshape(w.shape[0], 1, 1, Kd), None, bias, None, relu, 1, 0)
    return y.reshape(N, w.shape[0])


class _UpsampleAdd(torch.autograd.Function):
    """FPN top-down merge: lateral + nearest_x2(top)  (reference modeling/backbone/fpn.py:57-62)."""

    @staticmethod
    def forward(ctx, lateral, top):
        ctx.top_shape = top.shape
        return K.upsample2_add(lateral, top)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        g_top = None
        if ctx.needs_input_grad[1]:
            g_top = torch.zeros(ctx.top_shape, dtype=g.dtype, device=g.device)
            K.upsample2_backward_(g, g_top)
        return (g if ctx.needs_input_grad[0] else None), g_top


upsample_add = _UpsampleAdd.apply


class _Subsample2(torch.autograd.Function):
    """P6 = max_pool2d(P5, 1, 2, 0) = every other pixel (reference modeling/backbone/fpn.py:72-74)."""

    @staticmethod
    def forward(ctx, x):
        ctx.in_shape = x.shape
        return K.subsample2(x)

    @staticmethod
    def backward(ctx, g):
        gx = torch.zeros(ctx.in_shape, dtype=g.dtype, device=g.device)
        K.subsample2_backward_(g.contiguous(), gx)
        return gx


subsample2 = _Subsample2.apply


class _FpnRoiAlign(torch.autograd.Function):
    """All FPN levels in one launch (reference modeling/poolers.py:91-121 + layers/roi_align.py:11-44)."""

    @staticmethod
    def forward(ctx, rois, levels, scales, pooled, sampling_ratio, *feats):
        ctx.scales, ctx.pooled, ctx.sr = scales, pooled, sampling_ratio
        ctx.shapes = [f.shape for f in feats]
        ctx.save_for_backward(rois, levels)
        return K.roi_align_fpn_forward(list(feats), scales, rois, levels, pooled, sampling_ratio)

    @staticmethod
    def backward(ctx, g):
        rois, levels = ctx.saved_tensors
        grads = [torch.zeros(s, dtype=g.dtype, device=g.device) for s in ctx.shapes]
        K.roi_align_fpn_backward(g.contiguous(), ctx.scales, rois, levels, ctx.pooled, ctx.sr, grads)
        return (None, None, None, None, None) + tuple(grads)


def fpn_roi_align(feats, scales, rois, levels, pooled, sampling_ratio):
    return _FpnRoiAlign.apply(rois, levels, tuple(scales), pooled, sampling_ratio, *feats)


class _RpnLoss(torch.autograd.Function):
    """RPN objectness + box-regression losses (reference rpn/loss.py:138-196) over the fused head maps of all levels:
    one launch forward, one launch backward (analytic gradient written in the maps' own layout)."""

    @staticmethod
    def forward(ctx, labels, targets, pos, neg, n_sampled, num_anchors, beta, *fused):
        sums = K.rpn_loss_forward(fused, num_anchors, labels, targets, pos, neg, beta)
        ctx.save_for_backward(labels, targets, pos, neg, n_sampled, *fused)
        ctx.num_anchors, ctx.beta = num_anchors, beta
        out = sums / n_sampled
        return out[0], out[1]

    @staticmethod
    def backward(ctx, g_obj, g_box):
        labels, targets, pos, neg, n_sampled = ctx.saved_tensors[:5]
        fused = ctx.saved_tensors[5:]
        zero = n_sampled.new_zeros(())
        g_obj = zero if g_obj is None else g_obj.contiguous()
        g_box = zero if g_box is None else g_box.contiguous()
        dfused = K.rpn_loss_backward(fused, ctx.num_anchors, labels, targets, pos, neg, ctx.beta, n_sampled, g_obj, g_box)
        return (None,) * 7 + tuple(dfused)


def rpn_loss(fused, num_anchors, labels, targets, pos, neg, beta=1.0 / 9):
    """Returns (objectness_loss, box_loss).  labels fp32 / targets [.,4] fp32 / pos, neg bool, all over the batch's
    anchors in (image, level, pixel, anchor) order."""
    n_sampled = (pos | neg).sum().to(torch.float32)
    return _RpnLoss.apply(labels, targets, pos, neg, n_sampled, num_anchors, beta, *fused)


class _PsmLoss(torch.autograd.Function):
    """PSM consistency loss (reference box_head/loss.py:164-237,267-287,311-315) — one launch; the analytic
    gradient w.r.t. the student logits is produced by the same kernel."""

    @staticmethod
    def forward(ctx, student_logits, teacher_logits, labels, temperature, sharpen, hard_negative, neg_weight):
        loss, grad, sel = K.psm_loss(student_logits, teacher_logits, labels, temperature, sharpen, hard_negative,
                                     neg_weight)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(sel)
        return loss.reshape(()), sel

    @staticmethod
    def backward(ctx, g, _):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None, None, None


def psm_loss(student_logits, teacher_logits, labels, temperature, sharpen, hard_negative, neg_weight):
    return _PsmLoss.apply(student_logits, teacher_logits, labels, temperature, sharpen, hard_negative, neg_weight)


class _MgdLoss(torch.autograd.Function):
    """MGD masked feature-distillation loss (reference generalized_rcnn.py:243-282) for one student view:
    per level one pass reads the student map once and all teacher views, emitting the partial sums and the
    student gradient."""

    @staticmethod
    def forward(ctx, integral_mask, flip_mask, n_views, *tensors):
        n_levels = len(tensors) // (n_views + 1)
        students = tensors[:n_levels]
        teachers = tensors[n_levels:]  # view-major: teachers[v * n_levels + l]
        sizes = [(s.shape[1], s.shape[2]) for s in students]
        masks, sums = K.mgd_masks(integral_mask, sizes)
        C = students[0].shape[3]
        partials = torch.zeros((n_levels, n_views), dtype=torch.float32, device=students[0].device)
        want = any(ctx.needs_input_grad[3:3 + n_levels])
        grads = []
        w = 1.0 / (n_levels * n_views)
        for l in range(n_levels):
            tv = [teachers[v * n_levels + l] for v in range(n_views)]
            grads.append(K.mgd_level(students[l], tv, flip_mask, masks[l], sums[l:l + 1], w, partials[l], want))
        loss = K.mgd_finalize(partials, sums, n_levels, n_views, C)
        if want:
            ctx.save_for_backward(*grads)
        ctx.n_levels, ctx.n_rest = n_levels, len(teachers)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        grads = ctx.saved_tensors
        return (None, None, None) + tuple(gr * g for gr in grads) + (None,) * ctx.n_rest


def mgd_loss(integral_mask, student_levels, teacher_views, flip_mask):
    """student_levels: list of NHWC maps; teacher_views: list (views) of lists (levels) of NHWC maps (no grad)."""
    flat = [t for view in teacher_views for t in view]
    return _MgdLoss.apply(integral_mask, flip_mask, len(teacher_views), *student_levels, *flat)
