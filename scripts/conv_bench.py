"""Device-time of the tcgen05 conv kernels on the BASELINE shapes (B=2, 800x800) — CUDA events, L2 flushed
between iterations.  Prints TFLOP/s (2*MAC) per layer class; cuDNN (torch) TF32 shown beside for context."""
import os
import sys

import torch
import torch.nn.functional as Fn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn

dev = torch.device("cuda:0")
SHAPES = [  # name, B, H, W, Cin, Cout, k, stride, pad
    ("fpn_layer1 3x3 256->256 @200", 2, 200, 200, 256, 256, 3, 1, 1),
    ("fpn_layer2 3x3 256->256 @100", 2, 100, 100, 256, 256, 3, 1, 1),
    ("layer1 1x1 64->256 @200", 2, 200, 200, 64, 256, 1, 1, 0),
    ("layer1 3x3 64->64 @200", 2, 200, 200, 64, 64, 3, 1, 1),
    ("layer2 3x3 128->128 @100", 2, 100, 100, 128, 128, 3, 1, 1),
    ("layer3 3x3 256->256 @50", 2, 50, 50, 256, 256, 3, 1, 1),
    ("layer3 1x1 1024->256 @50", 2, 50, 50, 1024, 256, 1, 1, 0),
    ("layer4 3x3 512->512 @25", 2, 25, 25, 512, 512, 3, 1, 1),
    ("layer3.0 ds 1x1s2 512->1024", 2, 100, 100, 512, 1024, 1, 2, 0),
    ("rpn head 1x1 256->32 @200", 2, 200, 200, 256, 32, 1, 1, 0),
    ("mask fcn 3x3 256->256 @14 x256roi", 256, 14, 14, 256, 256, 3, 1, 1),
    ("fc6 12544->1024 x1024roi", 1, 1, 1024, 12544, 1024, 1, 1, 0),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
for name, B, H, W, Ci, Co, k, s, p in SHAPES:
    shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
    x = torch.randn(B, H, W, Ci, device=dev)
    w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
    dy = torch.randn(B, shape.out_h, shape.out_w, Co, device=dev)
    sc = torch.rand(Co, device=dev)
    flops = 2.0 * B * shape.out_h * shape.out_w * Co * Ci * k * k
    res = {}
    res["fprop"] = timeit(lambda: kn.conv_fprop(x, w, shape, sc, sc, None, True, kn.CONV_TCGEN05))
    res["dgrad"] = timeit(lambda: kn.conv_dgrad(dy, w, shape, kn.CONV_TCGEN05))
    res["wgrad"] = timeit(lambda: kn.conv_wgrad(x, dy, shape, False, kn.CONV_TCGEN05))
    xc = x.permute(0, 3, 1, 2)  # channels_last view for cuDNN
    wc = w.permute(0, 3, 1, 2)
    res["cudnn"] = timeit(lambda: Fn.conv2d(xc, wc, None, s, p))
    print("%-36s %6.2f GF | " % (name, flops / 1e9) + " | ".join(
        "%s %.3f ms %6.1f TF/s" % (k2, v[0], flops / v[0] / 1e9) for k2, v in res.items()), flush=True)
