"""Device idle gaps inside free-running iterations: where the GPU waits for the host (torch.profiler timeline)."""
import os, sys, json, gzip
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.config.presets import mt_psm_cfg
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
from mmt_psm_b200.utils.synthetic_init import init_model
from torch.profiler import ProfilerActivity, profile

dev = torch.device("cuda:0")
cfg = mt_psm_cfg(start_mt=0)
torch.manual_seed(0)
student = build_detection_model(cfg, is_student=True).to(dev)
teacher = build_detection_model(cfg, is_teacher=True).to(dev)
init_model(student, 0)
teacher.load_state_dict(student.state_dict())
opt = make_optimizer(cfg, student, teacher)
sched = make_lr_scheduler(cfg, opt)
sched.last_epoch = 2000
tr = MTtrainer(student, teacher, None, opt, sched, None, None, 10 ** 9, cfg, cuda_graphs=True)
student.train(); teacher.eval()
im, tg, vw = make_batch(1234, 2, 800, 3, 20, 20)
im, tg, vw = im.to(dev), [t.to(dev) for t in tg], [v.to(dev) for v in vw]
for j in range(4):
    tr.train_step(2000 + j, im, tg, vw)
torch.cuda.synchronize()
N = 3
marks = []
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for j in range(N):
        with torch.profiler.record_function("ITER%d" % j):
            tr.train_step(2100 + j, im, tg, vw)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
tr_ = json.load(open("/tmp/trace.json"))
ev = tr_["traceEvents"]
kern = sorted([(e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")])
iters = sorted([(e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev if e.get("cat") in ("user_annotation", "cpu_op") and str(e.get("name", "")).startswith("ITER")])
cpu_ops = sorted([(e["ts"], e["ts"] + e["dur"], e["name"]) for e in ev if e.get("cat") in ("cpu_op", "cuda_runtime", "python_function", "user_annotation")])
print("iterations:", [(round((b - a) / 1e3, 2)) for a, b, _ in iters])
# use the middle iteration's window [start of ITER1, start of ITER2]
a0, a1 = iters[1][0], iters[2][0]
inwin = [k for k in kern if k[0] >= a0 and k[0] < a1]
busy, cur = 0.0, None
gaps = []
prev_end, prev_name = a0, "ITER start"
for s, e, n in inwin:
    if s > prev_end + 100:
        gaps.append((prev_end - a0, s - prev_end, prev_name, n))
    if e > prev_end:
        busy += e - max(s, prev_end)
        prev_end, prev_name = e, n
print("window %.2f ms, device busy %.2f ms, %d kernels" % ((a1 - a0) / 1e3, busy / 1e3, len(inwin)))
gaps.sort(key=lambda g: -g[1])
print("largest idle gaps (offset ms, gap us, after kernel -> before kernel, host op at gap start):")
for off, g, pn, nn in gaps[:25]:
    t = a0 + off + 1
    host = [n for s, e, n in cpu_ops if s <= t < e]
    host = [h for h in host if not h.startswith("ITER")]
    print("  +%6.2f ms  %6.0f us   %-38s -> %-38s | %s" % (off / 1e3, g, pn[:38], nn[:38], " > ".join(h[:28] for h in host[-3:])))
print("total idle in gaps > 100 us: %.2f ms" % (sum(g[1] for g in gaps) / 1e3))
