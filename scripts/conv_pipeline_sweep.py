"""What bounds the K loop of the tcgen05 conv kernel?  Per layer: device time vs the depth of the shared-memory ring
(latency-bound: time ~ 1 / stages; bandwidth-bound: flat) and vs the distance of the producer's L2 prefetch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn, lib

L = lib.lib()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=4):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts) * 1e3


LAYERS = [("l3 3x3 256->256 B4", 4, 50, 256, 256, 3, 1), ("l3 3x3 256->256 B8", 8, 50, 256, 256, 3, 1),
          ("l2 3x3 128->128 B8", 8, 100, 128, 128, 3, 1), ("fpn_layer2 3x3 B4", 4, 100, 256, 256, 3, 1),
          ("fpn_layer1 3x3 B4", 4, 200, 256, 256, 3, 1), ("l3 1x1 1024->256 B4", 4, 50, 1024, 256, 1, 1),
          ("l3 1x1 256->1024 B4", 4, 50, 256, 1024, 1, 1), ("l4 3x3 512->512 B4", 4, 25, 512, 512, 3, 1),
          ("l4 1x1 2048->512 B4", 4, 25, 2048, 512, 1, 1), ("l1 3x3 64->64 B8", 8, 200, 64, 64, 3, 1),
          ("l2 1x1 128->512 B8", 8, 100, 128, 512, 1, 1), ("rpn conv 3x3 @25 B2", 2, 25, 256, 256, 3, 1)]
print("%-24s %2s | stages: %7s %7s %7s %7s | prefetch: %7s %7s %7s %7s" % ("layer", "cg", "2", "3", "4", "max", "2", "4", "8", "16"))
for name, B, H, Ci, Co, k, s in LAYERS:
    shape = kn.conv_shape(B, H, H, Ci, Co, k, s, k // 2)
    x = torch.randn(B, H, H, Ci, device=dev)
    w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
    sc = torch.rand(Co, device=dev)
    f = lambda: kn.conv_fprop(x, w, shape, sc, sc, None, True, kn.CONV_TCGEN05)
    for cg in (1, 2):
        kn.conv_force_cta_group(cg)
        row = []
        for st in (2, 3, 4, 0):
            L.mmt_conv_set_pipeline(st, 0)
            row.append(timeit(f))
        for pf in (2, 4, 8, 16):
            L.mmt_conv_set_pipeline(0, pf)
            row.append(timeit(f))
        L.mmt_conv_set_pipeline(0, 0)
        print("%-24s %2d | stages: %7.1f %7.1f %7.1f %7.1f | prefetch: %7.1f %7.1f %7.1f %7.1f" % ((name, cg) + tuple(row)), flush=True)
kn.conv_force_cta_group(0)
