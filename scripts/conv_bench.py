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


def all_shapes():
    """Every tcgen05 conv of the MT iteration as (name, B, H, W, Ci, Co, k, s, p, n_fwd2, n_fwd8, n_bwd):
    launches per iteration at batch 2 with grad (student sup + student unl = 2 passes), batch 8 no-grad (teacher), and
    backward passes (2, except the frozen layer1)."""
    out = []
    def add(name, H, Ci, Co, k, s, count, frozen=False, teacher=True, passes=2):
        out.append((name, H, Ci, Co, k, s, count, frozen, teacher, passes))
    add("l1 1x1 64->256", 200, 64, 256, 1, 1, 4, True)
    add("l1 1x1 64->64", 200, 64, 64, 1, 1, 1, True)
    add("l1 3x3 64->64", 200, 64, 64, 3, 1, 3, True)
    add("l1 1x1 256->64", 200, 256, 64, 1, 1, 2, True)
    add("l2 ds 1x1s2 256->512", 200, 256, 512, 1, 2, 1)
    add("l2 1x1s2 256->128", 200, 256, 128, 1, 2, 1)
    add("l2 3x3 128->128", 100, 128, 128, 3, 1, 4)
    add("l2 1x1 128->512", 100, 128, 512, 1, 1, 4)
    add("l2 1x1 512->128", 100, 512, 128, 1, 1, 3)
    add("l3 ds 1x1s2 512->1024", 100, 512, 1024, 1, 2, 1)
    add("l3 1x1s2 512->256", 100, 512, 256, 1, 2, 1)
    add("l3 3x3 256->256", 50, 256, 256, 3, 1, 6)
    add("l3 1x1 256->1024", 50, 256, 1024, 1, 1, 6)
    add("l3 1x1 1024->256", 50, 1024, 256, 1, 1, 5)
    add("l4 ds 1x1s2 1024->2048", 50, 1024, 2048, 1, 2, 1)
    add("l4 1x1s2 1024->512", 50, 1024, 512, 1, 2, 1)
    add("l4 3x3 512->512", 25, 512, 512, 3, 1, 3)
    add("l4 1x1 512->2048", 25, 512, 2048, 1, 1, 3)
    add("l4 1x1 2048->512", 25, 2048, 512, 1, 1, 2)
    for i, (h, c) in enumerate(((200, 256), (100, 512), (50, 1024), (25, 2048)), 1):
        add("fpn_inner%d 1x1 %d->256" % (i, c), h, c, 256, 1, 1, 1)
        add("fpn_layer%d 3x3 256->256" % i, h, 256, 256, 3, 1, 1)
    for h in (200, 100, 50, 25, 13):
        add("rpn conv 3x3 @%d" % h, h, 256, 256, 3, 1, 1, passes=1)
        add("rpn heads 1x1 256->16 @%d" % h, h, 256, 16, 1, 1, 1, passes=1)
        add("adaptor 1x1 256->128 @%d" % h, h, 256, 128, 1, 1, 1, passes=1)
    return out


if len(sys.argv) > 1 and sys.argv[1] == "--all":
    tot = dict(f2=0.0, f8=0.0, d=0.0, w=0.0, hbm=0.0)
    print("%-28s %3s | %8s %8s | %8s %8s | %8s %8s | %8s | per-iteration ms (count x passes)" %
          ("layer", "cnt", "f B2 us", "TF/s", "f B8 us", "TF/s", "dgrad us", "wgrad us", "hbm B2 us"))
    for name, H, Ci, Co, k, s, count, frozen, teacher, passes in all_shapes():
        p = k // 2
        r = {}
        for B in (2, 8):
            shape = kn.conv_shape(B, H, H, Ci, Co, k, s, p)
            x = torch.randn(B, H, H, Ci, device=dev)
            w = torch.randn(Co, k, k, Ci, device=dev) * 0.05
            sc = torch.rand(Co, device=dev)
            res = torch.randn(B, shape.out_h, shape.out_w, Co, device=dev) if (k == 1 and Co >= 4 * Ci) else None
            r[B] = timeit(lambda: kn.conv_fprop(x, w, shape, sc, sc, res, True, kn.CONV_TCGEN05))[0]
            if B == 2:
                dy = torch.randn(B, shape.out_h, shape.out_w, Co, device=dev)
                r["d"] = timeit(lambda: kn.conv_dgrad(dy, w, shape, kn.CONV_TCGEN05))[0] if Ci % 32 == 0 and Co % 32 == 0 else float("nan")
                r["w"] = timeit(lambda: kn.conv_wgrad(x, dy, shape, False, kn.CONV_TCGEN05))[0]
                flops = 2.0 * B * shape.out_h * shape.out_w * Co * Ci * k * k
                hbm = 4.0 * (x.numel() + w.numel() + dy.numel() * (2 if res is not None else 1)) / 6.5e12 * 1e6
        f2 = count * passes * r[2]
        f8 = count * r[8] if teacher else 0.0
        d = 0.0 if frozen else count * passes * r["d"]
        wg = 0.0 if frozen else count * passes * r["w"]
        tot["f2"] += f2; tot["f8"] += f8; tot["d"] += d if d == d else 0.0; tot["w"] += wg; tot["hbm"] += count * passes * hbm / 1e3
        print("%-28s %3d | %8.1f %8.1f | %8.1f %8.1f | %8.1f %8.1f | %8.1f | f2 %.3f f8 %.3f d %.3f w %.3f" %
              (name, count, r[2] * 1e3, flops / r[2] / 1e9, r[8] * 1e3, 4 * flops / r[8] / 1e9, r["d"] * 1e3, r["w"] * 1e3, hbm,
               f2, f8, d, wg), flush=True)
    print("TOTAL per iteration: fprop B2 %.2f ms, fprop B8 %.2f ms, dgrad %.2f ms, wgrad %.2f ms (HBM floor of the B2 fprops %.2f ms)" %
          (tot["f2"], tot["f8"], tot["d"], tot["w"], tot["hbm"]))
    sys.exit(0)

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
