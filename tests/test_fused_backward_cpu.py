"""CPU suite: the HOST logic of the block-level fused backward (functional._Bottleneck) — which gradient goes through
which kernel call with which epilogue — checked against torch autograd over the plain layer-by-layer composition.
The CUDA primitives are replaced here by torch restatements of their documented contracts (include/mmtpsm_b200.h:
mmt_conv_fprop_f32, mmt_conv_dgrad_fused_f32, mmt_conv_wgrad_scaled_f32, mmt_bn_relu_backward_f32); the kernels themselves
are tested against the same contracts on the GPU (tests/test_ops_gpu.py::test_conv_fused_backward_ops)."""
import sys

import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT

sys.path.insert(0, ROOT)


def _nchw(t):
    return t.permute(0, 3, 1, 2)


def _nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def _oihw(w):
    return w.permute(0, 3, 1, 2)


class _Emu(object):
    """torch restatements of the C-ABI contracts (NHWC activations, OHWI filters)."""

    @staticmethod
    def conv_fprop(x, w, shape, scale=None, bias=None, residual=None, relu=False, impl=None, out=None):
        y = _nhwc(F.conv2d(_nchw(x), _oihw(w), None, shape.stride, shape.pad))
        if scale is not None:
            y = y * scale
        if bias is not None:
            y = y + bias
        if residual is not None:
            y = y + residual
        return torch.relu(y) if relu else y

    @staticmethod
    def conv_dgrad_fused(dy, w, shape, k_scale=None, aux=None, aux_mode=0):
        g = dy if k_scale is None else dy * k_scale
        dx = _nhwc(torch.nn.grad.conv2d_input((shape.batch, shape.in_c, shape.in_h, shape.in_w), _oihw(w), _nchw(g),
                                              shape.stride, shape.pad))
        if aux_mode == 1:
            dx = dx + aux
        elif aux_mode == 2:
            dx = torch.where(aux > 0, dx, torch.zeros_like(dx))
        return dx

    @staticmethod
    def conv_wgrad_scaled(x, dy, shape, row_scale=None, into=None):
        dw = torch.nn.grad.conv2d_weight(_nchw(x), (shape.out_c, shape.in_c, shape.k_h, shape.k_w), _nchw(dy),
                                         shape.stride, shape.pad).permute(0, 2, 3, 1)
        if row_scale is not None:
            dw = dw * row_scale.view(-1, 1, 1, 1)
        if into is not None:
            into += dw
            return None
        return dw.contiguous()

    @staticmethod
    def bn_relu_backward(grad_y, y, scale, relu):
        g = grad_y * (y > 0) if relu else grad_y
        return g * scale if scale is not None else g


@pytest.fixture
def emulated(monkeypatch):
    import mmt_psm_b200  # noqa: F401
    from mmt_psm_b200 import functional as Fk, kernels as K
    for name in ("conv_fprop", "conv_dgrad_fused", "conv_wgrad_scaled", "bn_relu_backward"):
        monkeypatch.setattr(K, name, getattr(_Emu, name))
    monkeypatch.setattr(Fk, "_fused_grad_sink", None)
    return Fk


def _reference_block(x, w1, w2, w3, wd, s1, b1, s2, b2, s3, b3, sd, bd, stride):
    bn = lambda t, s, b: t * s.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
    xn = _nchw(x)
    idt = xn if wd is None else bn(F.conv2d(xn, _oihw(wd), None, stride), sd, bd)
    o = torch.relu(bn(F.conv2d(xn, _oihw(w1), None, stride), s1, b1))
    o = torch.relu(bn(F.conv2d(o, _oihw(w2), None, 1, 1), s2, b2))
    return _nhwc(torch.relu(bn(F.conv2d(o, _oihw(w3), None), s3, b3) + idt))


@pytest.mark.parametrize("downsample,stride,x_grad", [(False, 1, True), (True, 2, True), (True, 1, True), (True, 2, False)])
def test_bottleneck_fused_backward_equals_autograd(emulated, downsample, stride, x_grad):
    Fk = emulated
    g = torch.Generator().manual_seed(5)
    B, H, W, Cmid = 2, 10, 12, 8
    Cin = 16 if downsample else 32
    Cout = 32
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x = rnd(B, H, W, Cin)
    w1, w2, w3 = rnd(Cmid, 1, 1, Cin) * 0.3, rnd(Cmid, 3, 3, Cmid) * 0.2, rnd(Cout, 1, 1, Cmid) * 0.3
    wd = rnd(Cout, 1, 1, Cin) * 0.3 if downsample else None
    aff = lambda n: (torch.rand(n, generator=g, dtype=torch.float64) + 0.5, rnd(n) * 0.1)
    (s1, b1), (s2, b2), (s3, b3) = aff(Cmid), aff(Cmid), aff(Cout)
    sd, bd = aff(Cout) if downsample else (None, None)
    gy = None
    results = []
    for fn in (Fk.bottleneck, _reference_block):
        leaves = [t.clone().requires_grad_(True) for t in (x, w1, w2, w3)]
        leaves[0].requires_grad_(x_grad)
        wdl = wd.clone().requires_grad_(True) if downsample else None
        y = fn(leaves[0], leaves[1], leaves[2], leaves[3], wdl, s1, b1, s2, b2, s3, b3, sd, bd, stride)
        if gy is None:
            gy = rnd(*y.shape)
        y.backward(gy)
        results.append((y.detach(), [l.grad for l in leaves] + [wdl.grad if downsample else None]))
    (y0, g0), (y1, g1) = results
    assert torch.allclose(y0, y1, rtol=1e-12, atol=1e-12)
    for a, b, name in zip(g0, g1, ("x", "w1", "w2", "w3", "wd")):
        if b is None:
            assert a is None, name
            continue
        assert a is not None and a.shape == b.shape, name
        assert torch.allclose(a, b, rtol=1e-10, atol=1e-10), (name, float((a - b).abs().max()))


def test_bottleneck_fused_backward_accumulates_into_arena(emulated, monkeypatch):
    """With the optimizer's gradient sink installed the weight gradients are ADDED into the live `.grad` buffers and
    autograd receives None for them (no AccumulateGrad pass); the touched parameters are reported."""
    Fk = emulated
    g = torch.Generator().manual_seed(6)
    rnd = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x = rnd(1, 6, 6, 16).requires_grad_(True)
    ws = [torch.nn.Parameter(t) for t in (rnd(4, 1, 1, 16), rnd(4, 3, 3, 4), rnd(16, 1, 1, 4))]
    for w in ws:
        w.grad = torch.ones_like(w)
    one = lambda n: (torch.ones(n, dtype=torch.float64) * 1.5, torch.zeros(n, dtype=torch.float64))
    seen = []
    monkeypatch.setattr(Fk, "_fused_grad_sink", lambda ps: seen.extend(ps))
    y = Fk.bottleneck(x, ws[0], ws[1], ws[2], None, *one(4), *one(4), *one(16), None, None, 1)
    y.sum().backward()
    assert [id(p) for p in seen] == [id(ws[2]), id(ws[1]), id(ws[0])]
    ref = [torch.nn.Parameter(w.detach().clone()) for w in ws]
    xr = x.detach().clone().requires_grad_(True)
    _reference_block(xr, ref[0], ref[1], ref[2], None, *one(4), *one(4), *one(16), None, None, 1).sum().backward()
    for w, r in zip(ws, ref):
        assert torch.allclose(w.grad, r.grad + 1.0, rtol=1e-10, atol=1e-10)
    assert torch.allclose(x.grad, xr.grad, rtol=1e-10, atol=1e-10)
