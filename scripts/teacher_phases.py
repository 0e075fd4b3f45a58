"""Host-enqueue time vs completed time of the phases of the teacher pass (steady state, CUDA graphs on, sequential)."""
import os
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
im, tg, vw = make_batch(1234, 2, 800, 3, 20, 20)
im, tg, vw = im.to(dev), [t.to(dev) for t in tg], [v.to(dev) for v in vw]
for j in range(5):
    tr.train_step(2000 + j, im, tg, vw)
torch.cuda.synchronize()

acc = {}


def wrap(obj, name, label):
    orig = getattr(obj, name)

    def f(*a, **k):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = orig(*a, **k)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        e = acc.setdefault(label, [0.0, 0.0, 0])
        e[0] += t1 - t0; e[1] += t2 - t0; e[2] += 1
        return out
    setattr(obj, name, f)


wrap(teacher, "extract_aug_feat", "trunk (graph replay + split)")
wrap(teacher.rpn, "forward", "rpn eval (coarse proposals)")
wrap(teacher.box_heads, "forward", "box head eval + detections")
wrap(teacher.mask_heads, "forward", "mask head eval")
wrap(teacher.mask_heads.mask.mask_generator.masker, "integral_mask", "paste -> integral mask")
wrap(teacher.rpn, "forward_teacher", "rpn train-mode proposals")
wrap(teacher.box_heads, "forward_teacher", "subsample + 4-view box head")
N = 5
with torch.no_grad():
    for j in range(N):
        teacher.forward_teacher([vw[0], vw[1]])
print("%-32s %10s %12s" % ("teacher phase", "enqueue ms", "completed ms"))
for k, (a, b, n) in acc.items():
    print("%-32s %10.2f %12.2f" % (k, a / N * 1e3, b / N * 1e3))
print("%-32s %10.2f %12.2f" % ("SUM", sum(v[0] for v in acc.values()) / N * 1e3, sum(v[1] for v in acc.values()) / N * 1e3))
