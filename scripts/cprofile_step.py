"""cProfile of the host side of the iteration phases (which Python glue the enqueue time goes to)."""
import cProfile, io, os, pstats, sys
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
for j in range(3):
    tr.train_step(2000 + j, im, tg, vw)
torch.cuda.synchronize()
N = 3


def report(pr, title, n=45):
    s = io.StringIO()
    pstats.Stats(pr, stream=s).strip_dirs().sort_stats("cumulative").print_stats(n)
    print("=" * 20, title)
    print("\n".join(l[:160] for l in s.getvalue().splitlines()[4:]))


pr_t, pr_s, pr_u, pr_b = cProfile.Profile(), cProfile.Profile(), cProfile.Profile(), cProfile.Profile()
for j in range(N):
    pr_s.enable(); ld = student(im, tg); pr_s.disable()
    torch.cuda.synchronize()
    with torch.no_grad():
        pr_t.enable(); res = teacher.forward_teacher(vw[:2]); pr_t.disable()
    torch.cuda.synchronize()
    pr_u.enable(); ld.update(student.forward_student(vw[-1:], res)); pr_u.disable()
    torch.cuda.synchronize()
    losses = sum(v for v in tr.weight_sum_loss(ld, 2100 + j).values())
    opt.zero_grad()
    pr_b.enable(); losses.backward(); pr_b.disable()
    opt.step(ema_alpha=0.99)
    torch.cuda.synchronize()
report(pr_t, "teacher.forward_teacher x%d" % N, 60)
report(pr_s, "student.forward (supervised) x%d" % N, 60)
report(pr_u, "student.forward_student x%d" % N, 25)
report(pr_b, "backward x%d" % N, 25)
