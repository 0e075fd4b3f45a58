"""In-process A/B of the engine's knobs on the bench workload (MT iteration, 2 + 2 crops x 3 views, 800x800): the same
trainer, the same resident batches; for every configuration the CUDA graphs are re-captured, 4 iterations warmed up
and `--steps` iterations timed with per-step CUDA events (median and mean reported).  Configurations are visited twice
(A B C A B C) so that drift shows up as a difference between the two visits of the same configuration."""
import argparse
import os
import statistics
import sys

os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import mmt_psm_b200  # noqa: E402,F401
from mmt_psm_b200 import lib  # noqa: E402
from mmt_psm_b200.config.presets import mt_psm_cfg  # noqa: E402
from mmt_psm_b200.data.synthetic import make_batch  # noqa: E402
from mmt_psm_b200.engine import MTtrainer  # noqa: E402
from mmt_psm_b200.modeling.backbone import BottleneckWithFixedBatchNorm  # noqa: E402
from mmt_psm_b200.modeling.detector import build_detection_model  # noqa: E402
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer  # noqa: E402
from mmt_psm_b200.utils import streams  # noqa: E402
from mmt_psm_b200.utils.synthetic_init import init_model  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--configs", default="base,nopdl,cg1,unfused")
args = ap.parse_args()

dev = torch.device("cuda:0")
L = lib.lib()
cfg = mt_psm_cfg(start_mt=0)
torch.manual_seed(0)
student = build_detection_model(cfg, is_student=True).to(dev)
teacher = build_detection_model(cfg, is_teacher=True).to(dev)
init_model(student, 0)
teacher.load_state_dict(student.state_dict())
opt = make_optimizer(cfg, student, teacher)
sched = make_lr_scheduler(cfg, opt)
sched.last_epoch = 2000
trainer = MTtrainer(student, teacher, None, opt, sched, None, None, 10 ** 9, cfg, cuda_graphs=True)
student.train()
teacher.eval()
torch.manual_seed(100)
batches = []
for j in range(2):
    images, targets, views = make_batch(1234 + 1000 * j, 2, 800, 3, 20, 20)
    batches.append((images.to(dev), [t.to(dev) for t in targets], [v.to(dev) for v in views]))

KNOBS = {
    "base": dict(pdl=1, cg=0, fused=True),
    "nopdl": dict(pdl=0, cg=0, fused=True),
    "cg1": dict(pdl=1, cg=1, fused=True),
    "unfused": dict(pdl=1, cg=0, fused=False),
    "nothread": dict(pdl=1, cg=0, fused=True, thread=False),
    "nolevels": dict(pdl=1, cg=0, fused=True, levels=False),
    "teacherfirst": dict(pdl=1, cg=0, fused=True, teacher_first=True),
    "noshare": dict(pdl=1, cg=0, fused=True, rpn_share=False),
    "bwthread": dict(pdl=1, cg=0, fused=True, bw_thread=True),
    "lateTargets": dict(pdl=1, cg=0, fused=True, early_rpn=False),
    # head / glue work of both host threads on high-priority streams + teacher trunk captured with 132 of 148 CTAs are the
    # shipped defaults (MTtrainer.head_priority / teacher_trunk_ctas); these switch them off / vary them
    "noprio": dict(pdl=1, cg=0, fused=True, head_prio=0, tcap=148),
    "prio_only": dict(pdl=1, cg=0, fused=True, tcap=148),
    "tcap_only": dict(pdl=1, cg=0, fused=True, head_prio=0),
    "tcap116": dict(pdl=1, cg=0, fused=True, tcap=116),
    "tcap100": dict(pdl=1, cg=0, fused=True, tcap=100),
    # interpreter's thread switch interval (two host threads share the GIL): default 5 ms
    "sw1e4": dict(pdl=1, cg=0, fused=True, switch=1e-4),
}
it = [2000]


def run(n):
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    torch.cuda.synchronize()
    marks[0].record()
    for j in range(n):
        images, targets, views = batches[j % 2]
        trainer.train_step(it[0], images, targets, views)
        it[0] += 1
        marks[j + 1].record()
    torch.cuda.synchronize()
    return [marks[j].elapsed_time(marks[j + 1]) for j in range(n)]


names = [c for c in args.configs.split(",") if c in KNOBS]
for visit in range(2):
    for name in names:
        k = KNOBS[name]
        sys.setswitchinterval(k.get("switch", 0.005))
        torch.cuda.synchronize()
        trainer.head_priority = k.get("head_prio", -1)          # shipped default: -1 / 132 (MTtrainer.__init__)
        trainer.teacher_trunk_ctas = k.get("tcap", 132)
        trainer._pregrown_for = 0         # the per-stream allocator pools are grown again after this configuration's capture
        trainer.set_cuda_graphs(False)
        L.mmt_conv_set_pdl(k["pdl"])
        L.mmt_conv_force_cta_group(k["cg"])
        BottleneckWithFixedBatchNorm.fused_backward = k["fused"]
        trainer.teacher_thread = k.get("thread", True)
        trainer.backward_thread = k.get("bw_thread", False)
        trainer.early_rpn_targets = k.get("early_rpn", True)
        streams.set_enabled(k.get("levels", True))
        teacher.rpn.box_selector_train.shared["enabled"] = k.get("rpn_share", True)
        tf = k.get("teacher_first", False)
        trainer.teacher_first = tf
        teacher.side_priority = -1 if tf else 0
        teacher._side = None              # re-created with the configuration's priority
        trainer._t_stream = None
        trainer.set_cuda_graphs(True)
        run(4)
        t = run(args.steps)
        print("visit %d  %-8s median %.2f ms  mean %.2f  min %.2f  max %.2f" % (
            visit, name, statistics.median(t), sum(t) / len(t), min(t), max(t)), flush=True)
