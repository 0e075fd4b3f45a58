"""ResNet-50 + FPN backbone on the tcgen05 conv engine (NHWC activations).

Mirrors the module tree / parameter names of the reference (modeling/backbone/resnet.py:61-293,
fpn.py:7-74, backbone.py:12-44): `body.stem.{conv1,bn1}`, `body.layer{1..4}.{i}.{conv1,bn1,conv2,bn2,conv3,bn3,
downsample.0,downsample.1}`, `fpn.fpn_inner{1..4}`, `fpn.fpn_layer{1..4}`.  Every Conv -> FrozenBN -> ReLU
(-> +identity) chain of the reference is ONE kernel launch here (epilogue fusion)."""
from collections import OrderedDict, namedtuple

import torch
from torch import nn

from .. import functional as Fk
from .. import kernels as K
from ..layers import Conv2d, FrozenBatchNorm2d
from ..layers.conv import kaiming_uniform_ohwi_
from . import registry

StageSpec = namedtuple("StageSpec", ["index", "block_count", "return_features"])
_STAGE_SPECS = {
    "R-50-C4": tuple(StageSpec(i, c, r) for i, c, r in ((1, 3, False), (2, 4, False), (3, 6, True))),
    "R-50-C5": tuple(StageSpec(i, c, r) for i, c, r in ((1, 3, False), (2, 4, False), (3, 6, False), (4, 3, True))),
    "R-50-FPN": tuple(StageSpec(i, c, r) for i, c, r in ((1, 3, True), (2, 4, True), (3, 6, True), (4, 3, True))),
    "R-101-FPN": tuple(StageSpec(i, c, r) for i, c, r in ((1, 3, True), (2, 4, True), (3, 23, True), (4, 3, True))),
}


class BottleneckWithFixedBatchNorm(nn.Module):
    def __init__(self, in_channels, bottleneck_channels, out_channels, num_groups=1, stride_in_1x1=True, stride=1):
        super().__init__()
        if num_groups != 1:
            raise NotImplementedError("grouped 3x3 convolutions are not on the MMT-PSM path (NUM_GROUPS = 1)")
        self.downsample = None
        if in_channels != out_channels:
            self.downsample = nn.Sequential(Conv2d(in_channels, out_channels, 1, stride, 0, bias=False),
                                            FrozenBatchNorm2d(out_channels))
        s1, s3 = (stride, 1) if stride_in_1x1 else (1, stride)
        self.conv1 = Conv2d(in_channels, bottleneck_channels, 1, s1, 0, bias=False)
        self.bn1 = FrozenBatchNorm2d(bottleneck_channels)
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, 3, s3, 1, bias=False)
        self.bn2 = FrozenBatchNorm2d(bottleneck_channels)
        self.conv3 = Conv2d(bottleneck_channels, out_channels, 1, 1, 0, bias=False)
        self.bn3 = FrozenBatchNorm2d(out_channels)

    fused_backward = True   # one autograd node per block with the elementwise backward folded into the conv kernels

    def forward(self, x):
        if self.fused_backward and self.conv1.stride in (1, 2) and self.conv2.stride == 1:
            s1, b1 = self.bn1.affine()
            s2, b2 = self.bn2.affine()
            s3, b3 = self.bn3.affine()
            wd = sd = bd = None
            if self.downsample is not None:
                wd = self.downsample[0].weight
                sd, bd = self.downsample[1].affine()
            return Fk.bottleneck(x, self.conv1.weight, self.conv2.weight, self.conv3.weight, wd, s1, b1, s2, b2, s3, b3,
                                 sd, bd, self.conv1.stride)
        identity = x if self.downsample is None else self.downsample[0](x, bn=self.downsample[1])
        out = self.conv1(x, bn=self.bn1, relu=True)
        out = self.conv2(out, bn=self.bn2, relu=True)
        return self.conv3(out, bn=self.bn3, residual=identity, relu=True)


class StemWithFixedBatchNorm(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        out_channels = cfg.MODEL.RESNETS.STEM_OUT_CHANNELS
        self.conv1 = Conv2d(3, out_channels, 7, 2, 3, bias=False)
        self.bn1 = FrozenBatchNorm2d(out_channels)

    def takes_nchw(self, x_nchw):
        """True when the tensor-core stem applies: it then reads the reference's NCHW image batch directly."""
        c = self.conv1
        return (K.get_conv_math() == "tf32" and not c.weight.requires_grad and not x_nchw.requires_grad and x_nchw.is_cuda
                and (c.in_channels, c.kernel_size, c.stride, c.padding) == (3, 7, 2, 3) and x_nchw.shape[1] == 3)

    def forward(self, x, nchw=False):
        c = self.conv1
        if nchw or (self.takes_nchw(x.permute(0, 3, 1, 2)) if x.dim() == 4 else False):
            # frozen stem: forward only, on the tensor cores (space-to-depth implicit GEMM, csrc/conv_tc.cu)
            if nchw:
                B, _, H, W = x.shape
            else:
                B, H, W, _ = x.shape
            scale, shift = self.bn1.affine()
            shape = K.conv_shape(B, H, W, 3, c.out_channels, 7, 2, 3)
            x = K.conv_stem_fprop(x.contiguous(), c.weight.detach(), shape, scale, shift, relu=True, nchw=nchw)
        else:
            x = self.conv1(x, bn=self.bn1, relu=True)
        if x.requires_grad:
            raise NotImplementedError("the stem is frozen on this path (FREEZE_CONV_BODY_AT >= 1): no max-pool backward")
        return K.maxpool3x3s2(x)


_TRANSFORMATION_MODULES = {"BottleneckWithFixedBatchNorm": BottleneckWithFixedBatchNorm}
_STEM_MODULES = {"StemWithFixedBatchNorm": StemWithFixedBatchNorm}


class ResNet(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        r = cfg.MODEL.RESNETS
        self.stem = _STEM_MODULES[r.STEM_FUNC](cfg)
        block = _TRANSFORMATION_MODULES[r.TRANS_FUNC]
        in_channels = r.STEM_OUT_CHANNELS
        width = r.NUM_GROUPS * r.WIDTH_PER_GROUP
        self.stages, self.return_features = [], {}
        for spec in _STAGE_SPECS[cfg.MODEL.BACKBONE.CONV_BODY]:
            factor = 2 ** (spec.index - 1)
            out_channels = r.RES2_OUT_CHANNELS * factor
            blocks, stride = [], int(spec.index > 1) + 1
            for _ in range(spec.block_count):
                blocks.append(block(in_channels, width * factor, out_channels, r.NUM_GROUPS, r.STRIDE_IN_1X1, stride))
                stride, in_channels = 1, out_channels
            name = "layer%d" % spec.index
            self.add_module(name, nn.Sequential(*blocks))
            self.stages.append(name)
            self.return_features[name] = spec.return_features
        for idx in range(cfg.MODEL.BACKBONE.FREEZE_CONV_BODY_AT):
            m = self.stem if idx == 0 else getattr(self, "layer%d" % idx)
            for p in m.parameters():
                p.requires_grad = False

    def forward(self, x, nchw=False):
        outs = []
        x = self.stem(x, nchw=nchw)
        for name in self.stages:
            x = getattr(self, name)(x)
            if self.return_features[name]:
                outs.append(x)
        return outs


class LastLevelMaxPool(nn.Module):
    def forward(self, x):
        return [Fk.subsample2(x)]


class FPN(nn.Module):
    def __init__(self, in_channels_list, out_channels, top_blocks=None):
        super().__init__()
        self.inner_blocks, self.layer_blocks = [], []
        for idx, c in enumerate(in_channels_list, 1):
            inner, layer = Conv2d(c, out_channels, 1, 1, 0), Conv2d(out_channels, out_channels, 3, 1, 1)
            for m in (inner, layer):
                kaiming_uniform_ohwi_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)
            self.add_module("fpn_inner%d" % idx, inner)
            self.add_module("fpn_layer%d" % idx, layer)
            self.inner_blocks.append("fpn_inner%d" % idx)
            self.layer_blocks.append("fpn_layer%d" % idx)
        self.top_blocks = top_blocks

    def forward(self, x):
        """Reference fpn.py:43-69.  The four lateral 1x1 convs are independent of each other, and so are the four 3x3
        output convs once the top-down sums exist: each group is issued on parallel streams (utils/streams.py)."""
        from ..utils.streams import run_parallel
        dev = x[0].device
        call = lambda name, t: (lambda: getattr(self, name)(t))
        laterals = run_parallel([call(n, f) for n, f in zip(self.inner_blocks, x)], dev)
        inners = [None] * len(x)
        inners[-1] = laterals[-1]
        for i in range(len(x) - 2, -1, -1):
            inners[i] = Fk.upsample_add(laterals[i], inners[i + 1])
        results = list(run_parallel([call(n, t) for n, t in zip(self.layer_blocks, inners)], dev))
        if self.top_blocks is not None:
            results.extend(self.top_blocks(results[-1]))
        return tuple(results)


class _Backbone(nn.Sequential):
    """body (+ fpn).  Accepts the reference's NCHW image batch and converts once to NHWC."""

    def forward(self, images_nchw):
        mods = list(self)
        stem = getattr(mods[0], "stem", None)
        if stem is not None and hasattr(stem, "takes_nchw") and stem.takes_nchw(images_nchw):
            x = mods[0](images_nchw.float(), nchw=True)     # the tensor-core stem reads the NCHW batch as is
        else:
            x = mods[0](K.nchw_to_nhwc(images_nchw))
        for m in mods[1:]:
            x = m(x)
        return x


@registry.BACKBONES.register("R-50-C4")
def build_resnet_backbone(cfg):
    return _Backbone(OrderedDict([("body", ResNet(cfg))]))


@registry.BACKBONES.register("R-50-FPN")
@registry.BACKBONES.register("R-101-FPN")
def build_resnet_fpn_backbone(cfg):
    c2 = cfg.MODEL.RESNETS.RES2_OUT_CHANNELS
    fpn = FPN([c2, c2 * 2, c2 * 4, c2 * 8], cfg.MODEL.BACKBONE.OUT_CHANNELS, top_blocks=LastLevelMaxPool())
    return _Backbone(OrderedDict([("body", ResNet(cfg)), ("fpn", fpn)]))


def build_backbone(cfg):
    if cfg.MODEL.BACKBONE.CONV_BODY not in registry.BACKBONES:
        raise KeyError("cfg.MODEL.BACKBONE.CONV_BODY: %s is not registered" % cfg.MODEL.BACKBONE.CONV_BODY)
    return registry.BACKBONES[cfg.MODEL.BACKBONE.CONV_BODY](cfg)
