"""Which host thread is the longer one?  At the point where the main thread needs the teacher's results: is the future
already done (supervised thread longer) or how long does the main thread wait (teacher thread longer)?"""
import os
import statistics
import sys
import time

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.config.presets import mt_psm_cfg
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
from mmt_psm_b200.utils.synthetic_init import init_model

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
batches = []
for j in range(2):
    im, tg, vw = make_batch(1234 + 1000 * j, 2, 800, 3, 20, 20)
    batches.append((im.to(dev), [t.to(dev) for t in tg], [v.to(dev) for v in vw]))

rec = []
orig = tr.forward_unlabel
t_start = [0.0]


def probe(data_u_list=None, teacher_future=None):
    t0 = time.perf_counter()
    done = teacher_future.done() if teacher_future is not None else None
    if teacher_future is not None:
        teacher_future.result()
    t1 = time.perf_counter()
    out = orig(data_u_list, teacher_future)
    t2 = time.perf_counter()
    rec.append((done, (t0 - t_start[0]) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3))
    return out


tr.forward_unlabel = probe
for j in range(30):
    t_start[0] = time.perf_counter()
    tr.train_step(2000 + j, *batches[j % 2])
    if j >= 10:
        rec[-1] = rec[-1] + ((time.perf_counter() - t_start[0]) * 1e3,)
torch.cuda.synchronize()
r = [x for x in rec if len(x) == 5]
print("teacher already done at the join: %d of %d iterations" % (sum(1 for x in r if x[0]), len(r)))
print("main thread reaches the join at %.2f ms; waits %.2f ms for the teacher; student-unl forward (host) %.2f ms; train_step returns at %.2f ms (medians)" % tuple(
    statistics.median(x[i] for x in r) for i in (1, 2, 3, 4)))
