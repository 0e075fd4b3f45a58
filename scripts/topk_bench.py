"""Device time of the fused RPN top-k + gather (all 10 (image, level) problems of a pass, one launch) against torch.topk per
level + the gather kernel, on sigmoid scores distributed like an RPN's (most anchors near 0)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn

dev = torch.device("cuda:0")
A, Cp, N = 3, 16, 2
side = [200, 100, 50, 25, 13]
hw = [x * x for x in side]
g = torch.Generator(device="cuda").manual_seed(1)
scores = [torch.sigmoid(torch.randn(N, p * A, device=dev, generator=g) * 1.5 - 4.0) for p in hw]
fused = [torch.randn(N, x, x, Cp, device=dev, generator=g) for x in side]
anchors = [torch.rand(p * A, 4, device=dev, generator=g) for p in hw]
ks = [min(2000, p * A) for p in hw]
K_ = sum(ks)
d, a, s = torch.empty(N, K_, 4, device=dev), torch.empty(N, K_, 4, device=dev), torch.empty(N, K_, device=dev)


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def torch_path():
    col = 0
    for l, k in enumerate(ks):
        v, i = scores[l].topk(k, dim=1, sorted=True)
        kn.rpn_gather(fused[l], A, i, v, anchors[l], d, a, s, col)
        col += k


print("fused rpn_topk_gather: %.1f us   torch.topk x5 + gather x5: %.1f us" % (
    timeit(lambda: kn.rpn_topk_gather(scores, fused, A, anchors, ks, d, a, s)), timeit(torch_path)))
