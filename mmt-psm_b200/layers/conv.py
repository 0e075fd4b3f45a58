"""Conv2d / FrozenBatchNorm2d / ConvTranspose2d(2x2,s2) on the tcgen05 conv engine.

Same constructor arguments and parameter names as the reference's wrappers (layers/misc.py:30-64,
layers/batch_norm.py:6-24) so registry-built models keep their state-dict keys; internally the weight is
stored in the kernel layout OHWI ([out_c, kh, kw, in_c], fp32) and activations are NHWC.  The reference
layout (OIHW / IOHW) is produced / consumed only at the state-dict boundary (`reference_weight`,
`load_reference_weight`; see modeling/detector.py `reference_state_dict`)."""
import math

import torch
from torch import nn

from .. import functional as Fk


def kaiming_uniform_ohwi_(w, a=1.0):
    """nn.init.kaiming_uniform_(w_oihw, a) for a weight stored OHWI (fan_in = in_c * kh * kw)."""
    fan_in = w.shape[1] * w.shape[2] * w.shape[3] if w.dim() == 4 else w.shape[1]
    bound = math.sqrt(6.0 / ((1 + a * a) * fan_in))
    with torch.no_grad():
        return w.uniform_(-bound, bound)


def kaiming_normal_fanout_ohwi_(w, fan_out=None):
    """nn.init.kaiming_normal_(w_oihw, mode="fan_out", nonlinearity="relu") for OHWI storage."""
    if fan_out is None:
        fan_out = w.shape[0] * w.shape[1] * w.shape[2]
    with torch.no_grad():
        return w.normal_(0, math.sqrt(2.0 / fan_out))


class FrozenBatchNorm2d(nn.Module):
    """Fixed batch statistics and affine parameters, folded into the producing conv's epilogue:
    scale = weight * rsqrt(running_var) (no eps, as the reference), shift = bias - running_mean * scale."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))
        self._cache = None

    def affine(self):
        key = (self.weight._version, self.bias._version, self.running_mean._version, self.running_var._version,
               self.weight.data_ptr())
        if self._cache is None or self._cache[0] != key:
            scale = self.weight * self.running_var.rsqrt()
            shift = self.bias - self.running_mean * scale
            self._cache = (key, scale.contiguous(), shift.contiguous())
        return self._cache[1], self._cache[2]


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding
        self.weight = nn.Parameter(torch.empty(out_channels, kernel_size, kernel_size, in_channels))  # OHWI
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):  # nn.Conv2d default (kaiming_uniform, a = sqrt(5))
        fan_in = self.in_channels * self.kernel_size * self.kernel_size
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def forward(self, x, bn=None, residual=None, relu=False):
        """x NHWC.  `bn`: a FrozenBatchNorm2d folded into the epilogue; `residual`: NHWC tensor added before relu."""
        scale, shift = (None, self.bias) if bn is None else bn.affine()
        return Fk.conv_bn_act(x, self.weight, scale, shift, residual, relu, self.stride, self.padding)

    # ---- state-dict boundary (reference layout OIHW)
    def reference_weight(self):
        return self.weight.detach().permute(0, 3, 1, 2).contiguous()

    def load_reference_weight(self, w_oihw):
        with torch.no_grad():
            self.weight.copy_(w_oihw.permute(0, 2, 3, 1))


class ConvTranspose2x2(nn.Module):
    """ConvTranspose2d(in, out, 2, 2, 0) as used by the mask predictor (reference roi_mask_predictors.py:24):
    out[2y+a, 2x+b] = W[:, :, a, b]^T x[y, x] + bias — four 1x1 GEMMs fused into ONE 1x1 conv with 4*out
    channels ordered (a, b, co), followed by a pixel shuffle.  weight: OHWI of that conv, [4*out, 1, 1, in]."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(4 * out_channels, 1, 1, in_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))

    def forward(self, x, relu=False):
        B, H, W, _ = x.shape
        Co = self.out_channels
        y = Fk.conv_bn_act(x, self.weight, None, self.bias.repeat(4), None, relu, 1, 0)  # [B,H,W,(a,b,co)]
        return y.view(B, H, W, 2, 2, Co).permute(0, 1, 3, 2, 4, 5).reshape(B, 2 * H, 2 * W, Co)

    def reference_weight(self):  # IOHW [in, out, 2, 2]
        w = self.weight.detach().view(2, 2, self.out_channels, self.in_channels)
        return w.permute(3, 2, 0, 1).contiguous()

    def load_reference_weight(self, w_iohw):
        with torch.no_grad():
            self.weight.copy_(w_iohw.permute(2, 3, 1, 0).reshape(4 * self.out_channels, 1, 1, self.in_channels))
