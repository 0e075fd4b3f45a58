"""Where one MT iteration spends its time: per-phase host enqueue time vs. device completion time (a phase whose
enqueue time ~ its total time is launch/CPU bound), then a torch.profiler kernel / CPU-op summary of full iterations.

    python scripts/profile_step.py [iters] [out_prefix]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.config.presets import mt_psm_cfg
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
from mmt_psm_b200.utils.synthetic_init import init_model

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
prefix = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/profile_step"
dev = torch.device("cuda:0")
if os.environ.get("MMT_CONV_MAX_CTAS"):
    from mmt_psm_b200 import lib as _lib
    _lib.check(_lib.lib().mmt_conv_set_max_ctas(int(os.environ["MMT_CONV_MAX_CTAS"])))
QUICK = os.environ.get("MMT_PROFILE_QUICK") == "1"
cfg = mt_psm_cfg(start_mt=0, relation=os.environ.get("MMT_RELATION") == "1")   # MMT_RELATION=1: BASELINE config 4
torch.manual_seed(0)
student = build_detection_model(cfg, is_student=True).to(dev)
teacher = build_detection_model(cfg, is_teacher=True).to(dev)
init_model(student, 0)
teacher.load_state_dict(student.state_dict())
opt = make_optimizer(cfg, student, teacher)
sched = make_lr_scheduler(cfg, opt)
sched.last_epoch = 2000
tr = MTtrainer(student, teacher, None, opt, sched, None, None, 10 ** 9, cfg, cuda_graphs=os.environ.get("MMT_GRAPHS", "1") == "1")
student.train(); teacher.eval()
pool = []
for j in range(2):
    im, tg, vw = make_batch(1234 + 1000 * j, 2, 800, 3, 20, 20)
    pool.append((im.to(dev), [t.to(dev) for t in tg], [v.to(dev) for v in vw]))


def phased_step(it, images, targets, views, acc):
    def mark(name, t0):
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        a = acc.setdefault(name, [0.0, 0.0])
        a[0] += t1 - t0
        a[1] += t2 - t0
        return time.perf_counter()

    t = time.perf_counter()
    ld = tr.student(images, targets)
    t = mark("student_sup_fwd", t)
    teacher_list = views[:tr.teacher_bs]
    with torch.no_grad():
        res = tr.teacher.forward_teacher(teacher_list)
    t = mark("teacher_fwd", t)
    ld.update(tr.student.forward_student(views[-tr.student_bs:], res))
    t = mark("student_unl_fwd", t)
    tr.scheduler.step()
    wl = tr.weight_sum_loss(ld, it)
    losses = sum(v for v in wl.values())
    tr.optimizer.zero_grad()
    t = mark("loss_sum_zero_grad", t)
    losses.backward()
    t = mark("backward", t)
    tr.optimizer.step(ema_alpha=0.99, grad_scale=1.0)
    t = mark("sgd_ema", t)
    return losses


for j in range(3):
    tr.train_step(2000 + j, *pool[j % 2])
torch.cuda.synchronize()

acc = {}
for j in range(iters):
    phased_step(2100 + j, *pool[j % 2], acc)
print("phase                      enqueue ms   total ms   (per iteration, sync at phase boundaries)")
tot = [0.0, 0.0]
for k, (a, b) in acc.items():
    print("%-26s %9.2f %10.2f" % (k, a / iters * 1e3, b / iters * 1e3))
    tot[0] += a; tot[1] += b
print("%-26s %9.2f %10.2f" % ("SUM", tot[0] / iters * 1e3, tot[1] / iters * 1e3))

# free-running iterations (no syncs): wall per step
torch.cuda.synchronize()
t0 = time.perf_counter()
for j in range(iters):
    tr.train_step(2200 + j, *pool[j % 2])
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("free-running: enqueue %.2f ms/step, wall %.2f ms/step" % ((t1 - t0) / iters * 1e3, (t2 - t0) / iters * 1e3))

if QUICK:
    sys.exit(0)
from torch.profiler import ProfilerActivity, profile

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for j in range(iters):
        tr.train_step(2300 + j, *pool[j % 2])
    torch.cuda.synchronize()
ka = prof.key_averages()
with open(prefix + "_cuda.txt", "w") as f:
    f.write(ka.table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=90))
with open(prefix + "_cpu.txt", "w") as f:
    f.write(ka.table(sort_by="self_cpu_time_total", row_limit=60, max_name_column_width=90))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
kern_us = sum(e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total for e in ev)
print("profiler: %d device events, %.2f ms device time per iteration" % (len(ev), kern_us / iters / 1e3))
# busy time = union of kernel intervals
iv = sorted((e.time_range.start, e.time_range.end) for e in ev)
busy, cur_s, cur_e = 0.0, None, None
for s, e in iv:
    if cur_e is None or s > cur_e:
        if cur_e is not None:
            busy += cur_e - cur_s
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
if cur_e is not None:
    busy += cur_e - cur_s
span = (iv[-1][1] - iv[0][0]) if iv else 0
print("profiler: device busy %.2f ms / iteration over a span of %.2f ms / iteration" % (busy / iters / 1e3, span / iters / 1e3))
