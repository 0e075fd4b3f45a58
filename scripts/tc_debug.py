"""Per-case error report of the tcgen05 conv kernels vs a float64 torch reference (GPU box only).
usage: python scripts/tc_debug.py fprop|dgrad|wgrad"""
import os
import sys

import torch
import torch.nn.functional as Fn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import kernels as kn
from oracle import ops_ref as R

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_ops_gpu import CONV_CASES

op = sys.argv[1]
dev = torch.device("cuda:0")
gen = torch.Generator().manual_seed(1)


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-20))


for case in CONV_CASES:
    B, H, W, Ci, Co, k, s, p = case
    x = torch.randn(B, Ci, H, W, generator=gen)
    w = torch.randn(Co, Ci, k, k, generator=gen) / (Ci * k * k) ** 0.5
    shape = kn.conv_shape(B, H, W, Ci, Co, k, s, p)
    dy = torch.randn(B, Co, shape.out_h, shape.out_w, generator=gen)
    xd = x.permute(0, 2, 3, 1).contiguous().to(dev)
    wd = w.permute(0, 2, 3, 1).contiguous().to(dev)
    dyd = dy.permute(0, 2, 3, 1).contiguous().to(dev)
    try:
        if op == "fprop":
            y = kn.conv_fprop(xd, wd, shape, None, None, None, False, kn.CONV_TCGEN05).permute(0, 3, 1, 2).cpu()
            ex = Fn.conv2d(x.double(), w.double(), None, s, p).float()
            tr = Fn.conv2d(R.tf32_round(x).double(), R.tf32_round(w).double(), None, s, p).float()
            nr = Fn.conv2d(R.tf32_round(x, "nearest").double(), R.tf32_round(w, "nearest").double(), None, s, p).float()
            print(op, case, "fp32 %.2e trunc %.2e nearest %.2e" % (rel(y, ex), rel(y, tr), rel(y, nr)), flush=True)
        else:
            xr = x.double().requires_grad_(True)
            wr = w.double().requires_grad_(True)
            Fn.conv2d(xr, wr, None, s, p).backward(dy.double())
            if op == "dgrad":
                dx = kn.conv_dgrad(dyd, wd, shape, kn.CONV_TCGEN05).permute(0, 3, 1, 2).cpu()
                print(op, case, "fp32 %.2e" % rel(dx, xr.grad.float()), flush=True)
            else:
                dw, db = kn.conv_wgrad(xd, dyd, shape, True, kn.CONV_TCGEN05)
                print(op, case, "fp32 %.2e  dbias %.2e" % (rel(dw.permute(0, 3, 1, 2).cpu(), wr.grad.float()),
                                                          rel(db.cpu(), dy.sum((0, 2, 3)))), flush=True)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(op, case, "FAILED:", repr(e)[:300], flush=True)
        break
