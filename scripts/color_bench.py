"""Device colour views vs the reference's host pipeline (PIL ops through torchvision.transforms.functional, the
reference's `transpart2`): time per unlabeled batch (B crops x V views at 800x800) and achieved HBM GB/s.
Algorithmic bytes per (view, image): 3 B/pixel read twice (luma pass + main pass) + 12 B/pixel written."""
import os, random, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.data.unlabeled_views import colour_views, draw_view_params

B, V, S = 2, 3, 800
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
imgs = rng.integers(0, 256, (B, S, S, 3), dtype=np.uint8)
random.seed(1); np.random.seed(1)
params = [[draw_view_params(S, S) for _ in range(B)] for _ in range(V)]
d = torch.from_numpy(imgs).to(dev)
for _ in range(3):
    colour_views(d, params)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record(); colour_views(d, params); b.record()
    torch.cuda.synchronize()
    ts.append((a.elapsed_time(b), (time.perf_counter() - t0) * 1e3))
dev_ms, wall_ms = min(t[0] for t in ts), min(t[1] for t in ts)
bytes_ = V * B * S * S * (3 + 3 + 12)
print("device: %.3f ms kernels (%.0f GB/s algorithmic), %.3f ms incl. host-side parameter upload" % (dev_ms, bytes_ / dev_ms / 1e6, wall_ms))
from PIL import Image
import torchvision.transforms.functional as tf
t0 = time.perf_counter()
for v in range(V):
    for b in range(B):
        p = params[v][b]
        im = Image.fromarray(imgs[b], "RGB")
        im = tf.adjust_hue(tf.adjust_contrast(tf.adjust_brightness(im, p["brightness"]), p["contrast"]), p["hue"])
        a = np.array(im)
        for top, left, fill in p["rects"]:
            a[top:top + fill.shape[0], left:left + fill.shape[1]] = fill
        t = tf.to_tensor(Image.fromarray(a, "RGB"))
        t = tf.normalize(t[[2, 1, 0]] * 255, mean=[102.9801, 115.9465, 122.7717], std=[1., 1., 1.])
cpu_ms = (time.perf_counter() - t0) * 1e3
print("reference host pipeline (PIL, 1 core): %.1f ms for the same %d views -> %.0fx" % (cpu_ms, V * B, cpu_ms / wall_ms))
