"""Per-layer device time of the tcgen05 fprop / dgrad kernel under the tile configurations of csrc/conv_tc.cu:
one CTA per 128 x BN tile (cta_group::1) vs CTA pairs on 256 x BN tiles (cta_group::2) vs the library's cost model.
Batches as the trainer runs them: student trunk [4], teacher trunk [8], RPN head on [2].  CUDA events, L2 flushed."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, iters=4):
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
    return min(ts)


LAYERS = [  # name, H, Ci, Co, k, s, count per pass, batches
    ("l1 1x1 64->256", 200, 64, 256, 1, 1, 4, (4, 8)), ("l1 3x3 64->64", 200, 64, 64, 3, 1, 3, (4, 8)),
    ("l1 1x1 256->64", 200, 256, 64, 1, 1, 2, (4, 8)),
    ("l2 ds 1x1s2 256->512", 200, 256, 512, 1, 2, 1, (4, 8)), ("l2 3x3 128->128", 100, 128, 128, 3, 1, 4, (4, 8)),
    ("l2 1x1 128->512", 100, 128, 512, 1, 1, 4, (4, 8)), ("l2 1x1 512->128", 100, 512, 128, 1, 1, 3, (4, 8)),
    ("l3 ds 1x1s2 512->1024", 100, 512, 1024, 1, 2, 1, (4, 8)), ("l3 3x3 256->256", 50, 256, 256, 3, 1, 6, (4, 8)),
    ("l3 1x1 256->1024", 50, 256, 1024, 1, 1, 6, (4, 8)), ("l3 1x1 1024->256", 50, 1024, 256, 1, 1, 5, (4, 8)),
    ("l4 ds 1x1s2 1024->2048", 50, 1024, 2048, 1, 2, 1, (4, 8)), ("l4 3x3 512->512", 25, 512, 512, 3, 1, 3, (4, 8)),
    ("l4 1x1 512->2048", 25, 512, 2048, 1, 1, 3, (4, 8)), ("l4 1x1 2048->512", 25, 2048, 512, 1, 1, 2, (4, 8)),
    ("fpn_inner1 1x1 256->256", 200, 256, 256, 1, 1, 1, (4, 8)), ("fpn_layer1 3x3", 200, 256, 256, 3, 1, 1, (4, 8)),
    ("fpn_layer2 3x3", 100, 256, 256, 3, 1, 1, (4, 8)), ("fpn_layer3 3x3", 50, 256, 256, 3, 1, 1, (4, 8)),
    ("fpn_inner4 1x1 2048->256", 25, 2048, 256, 1, 1, 1, (4, 8)),
    ("rpn conv 3x3 @200", 200, 256, 256, 3, 1, 1, (2,)), ("rpn conv 3x3 @100", 100, 256, 256, 3, 1, 1, (2,)),
    ("adaptor 1x1 256->128 @200", 200, 256, 128, 1, 1, 1, (2, 8)),
    ("mask fcn 3x3 @14 x256 roi", 14, 256, 256, 3, 1, 4, (256,)), ("fc6 12544->1024 x1024", None, 12544, 1024, 1, 1, 1, (1,)),
]

print("%-28s %4s | %9s %9s %9s | %7s %7s | dgrad: %9s %9s" % ("layer", "B", "cg1 us", "cg2 us", "auto us", "cg2/cg1", "TF/s", "cg1 us", "cg2 us"))
tot = {1: 0.0, 2: 0.0, 0: 0.0}
for name, H, Ci, Co, k, s, count, batches in LAYERS:
    for B in batches:
        if H is None:
            shape = kn.conv_shape(1, 1, 1024, Ci, Co, 1, 1, 0)
            x = torch.randn(1, 1, 1024, Ci, device=dev)
        else:
            shape = kn.conv_shape(B, H, H, Ci, Co, k, s, k // 2)
            x = torch.randn(B, H, H, Ci, device=dev)
        w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
        sc = torch.rand(Co, device=dev)
        dy = torch.randn(shape.batch, shape.out_h, shape.out_w, Co, device=dev)
        flops = 2.0 * shape.batch * shape.out_h * shape.out_w * Co * Ci * k * k
        t, d = {}, {}
        for cg in (1, 2, 0):
            kn.conv_force_cta_group(cg)
            t[cg] = timeit(lambda: kn.conv_fprop(x, w, shape, sc, sc, None, True, kn.CONV_TCGEN05))
            if cg and Co % 32 == 0:
                d[cg] = timeit(lambda: kn.conv_dgrad(dy, w, shape, kn.CONV_TCGEN05))
            tot[cg] += count * t[cg]
        kn.conv_force_cta_group(0)
        print("%-28s %4d | %9.1f %9.1f %9.1f | %7.2f %7.0f | dgrad: %9.1f %9.1f" % (
            name, B, t[1] * 1e3, t[2] * 1e3, t[0] * 1e3, t[2] / t[1], flops / min(t.values()) / 1e9,
            d.get(1, float("nan")) * 1e3, d.get(2, float("nan")) * 1e3), flush=True)
print("sum over (count x listed batches): cg1 %.2f ms, cg2 %.2f ms, auto %.2f ms" % (tot[1], tot[2], tot[0]))
