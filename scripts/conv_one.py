"""One conv shape, a few launches — the target of single-kernel `ncu --set full` captures.

    python scripts/conv_one.py B H W Ci Co k stride pad [res] [mode=fprop|dgrad|wgrad] [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn

B, H, W, Ci, Co, k, s, p = [int(a) for a in sys.argv[1:9]]
use_res = len(sys.argv) > 9 and sys.argv[9] == "res"
mode = sys.argv[10] if len(sys.argv) > 10 else "fprop"
iters = int(sys.argv[11]) if len(sys.argv) > 11 else 3
dev = torch.device("cuda:0")
shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
x = torch.randn(B, H, W, Ci, device=dev)
w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
sc = torch.rand(Co, device=dev)
dy = torch.randn(B, shape.out_h, shape.out_w, Co, device=dev)
res = torch.randn_like(dy) if use_res else None
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for i in range(iters):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    if mode == "fprop":
        kn.conv_fprop(x, w, shape, sc, sc, res, True, kn.CONV_TCGEN05)
    elif mode == "dgrad":
        kn.conv_dgrad(dy, w, shape, kn.CONV_TCGEN05)
    else:
        kn.conv_wgrad(x, dy, shape, False, kn.CONV_TCGEN05)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
flops = 2.0 * B * shape.out_h * shape.out_w * Co * Ci * k * k
print("%s %s: best %.1f us, %.1f TF/s" % (mode, sys.argv[1:9], min(ts) * 1e3, flops / min(ts) / 1e9))
