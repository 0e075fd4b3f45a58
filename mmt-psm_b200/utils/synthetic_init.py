"""Deterministic, variance-calibrated random weights keyed by state-dict name.

There is no network access for the ImageNet / COCO checkpoint the reference starts from
(`e2e_mask_rcnn_R_50_FPN_1x.pth`, tools/train_mean_teacher.py:37-43), and a ResNet with FROZEN batch norm and
default random conv weights explodes numerically.  `synthetic_state_dict` produces, for any {name: shape} table in
the REFERENCE's layout, the same tensors on every machine (torch CPU generator seeded per name), scaled so
activations stay O(1): He-normal convs, frozen-BN scale 1 on the bottleneck's inner convs and the downsample (skip) branch, 0.25 on the
residual-closing conv3, 1/64 on the stem (pixel range -> O(1)), and MGD adaptors at 0.1 x He (a trained student / teacher
pair maps the views to nearby embeddings; with unit-scale random adaptors the un-normalised L2 hint loss starts at O(1) and
SGD at the cfg learning rate diverges within 5 iterations — measured, profiles/r01_loss_trace_*.log), and the IR-Net
duplicate-removal head's ROI-feature embedding at 0.1 x (its IoU regressor is linear in the fc7 features, whose norm swings 2-8 x
during the first iterations of a randomly initialised box head: at full scale nms_loss goes 0.1 -> 35 -> NaN in four
iterations, at 0.1 x it decreases monotonically — profiles/r01_loss_trace_relation_cpu_v6.log).  Used by bench.py, smoke() and the step-level parity fixtures (the
reference model is loaded with exactly the same dict)."""
import math
import os
import zlib

import torch


def _gen(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) + 1000003 * seed) % (2 ** 31))


def synthetic_state_dict(shapes, seed=0, adaptor_scale=None, box_reg_gain=1.0):
    if adaptor_scale is None:
        adaptor_scale = float(os.environ.get("MMT_SYNTH_ADAPTOR_SCALE", "0.1"))
    out = {}
    for name, shape in shapes.items():
        g = _gen(name, seed)
        leaf = name.split(".")[-1]
        parent = name.split(".")[-2] if "." in name else ""
        if "cell_anchors" in name:
            continue
        if leaf in ("running_mean",):
            t = torch.randn(shape, generator=g) * 0.05
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) * 0.5 + 0.75
        elif leaf == "weight" and len(shape) == 1:      # frozen BN gamma
            base = 0.25 if parent == "bn3" else 1.0
            if ".stem." in name:
                base = 1.0 / 64  # brings the +-100 pixel range (BGR255 - mean) down to O(1) activations
            t = base * (torch.rand(shape, generator=g) * 0.2 + 0.9)
        elif leaf == "bias":
            t = torch.randn(shape, generator=g) * 0.01
            if name.endswith("relation_nms.classifier.bias"):
                t = t + 0.5                              # a trained IoU regressor answers ~0.5 on real objects (FG_THREAD 0.1)
        elif leaf == "weight" and len(shape) == 4:
            if ("conv5_mask" in name and "appearance_feature_extractor" not in name) or "deconv_1" in name:  # ConvTranspose IOHW
                fan_in = shape[0]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            std = math.sqrt(2.0 / fan_in)
            if any(k in name for k in ("cls_logits", "bbox_pred", "mask_fcn_logits")):
                std = 0.01
            if "hint_adaptor" in name:
                std *= adaptor_scale                     # trained adaptors map both networks to nearby embeddings
            if "fpn_layer" in name:
                std *= 0.25                              # pyramid outputs O(1): what the heads of a trained net see
            t = torch.randn(shape, generator=g) * std
        elif leaf == "weight" and len(shape) == 2:       # linear layers
            std = math.sqrt(2.0 / shape[1]) * 0.5          # fc6 / fc7: keep |fc7(x)|^2 ~ 1e2 so SGD at the cfg lr is stable
            if "cls_score" in name:
                std = 0.01
            if "bbox_pred" in name:
                std = 0.001 * box_reg_gain             # > 1: parity fixtures that must exercise the box-regression loss
            if "roi_feat_embedding_fc" in name:
                std *= float(os.environ.get("MMT_SYNTH_RELATION_SCALE", "0.1"))   # see the module docstring (IR-Net head)
            t = torch.randn(shape, generator=g) * std
        else:
            t = torch.randn(shape, generator=g) * 0.01
        out[name] = t.float()
    return out


def init_model(model, seed=0, **kw):
    """Fill `model` (this package's GeneralizedRCNN) through the reference-layout boundary."""
    from ..modeling.detector import load_reference_state_dict, reference_state_dict
    shapes = {k: tuple(v.shape) for k, v in reference_state_dict(model).items()}
    sd = synthetic_state_dict(shapes, seed, **kw)
    load_reference_state_dict(model, sd)
    return sd
