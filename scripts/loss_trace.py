"""Loss trajectory of the bench workload for N iterations (stability check of the synthetic init at the cfg LR)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.config.presets import mt_psm_cfg
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
from mmt_psm_b200.utils.synthetic_init import init_model

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
last_epoch = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
size = int(sys.argv[3]) if len(sys.argv) > 3 else 800
dev = torch.device("cuda:0")
cfg = mt_psm_cfg(start_mt=0)
torch.manual_seed(0)
student = build_detection_model(cfg, is_student=True).to(dev)
teacher = build_detection_model(cfg, is_teacher=True).to(dev)
init_model(student, 0)
teacher.load_state_dict(student.state_dict())
opt = make_optimizer(cfg, student, teacher)
sched = make_lr_scheduler(cfg, opt)
sched.last_epoch = last_epoch
trainer = MTtrainer(student, teacher, None, opt, sched, None, None, 10 ** 9, cfg)
student.train(); teacher.eval()
pool = [make_batch(1234 + 1000 * j, 2, size, 3, 20, 20) for j in range(2)]
for it in range(n):
    images, targets, views = pool[it % 2]
    ld, total = trainer.train_step(2000 + it, images, targets, views)
    print(it, "lr %.5f" % opt.lr, "total %.4f" % float(total), {k: round(float(v), 4) for k, v in ld.items()}, flush=True)
