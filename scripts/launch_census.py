"""Kernel launches (cudaLaunchKernel / memcpy / memset calls) per glue region of one iteration: where the host-paced
tiny launches come from.  Regions are wrapped with record_function by monkeypatching (no package changes)."""
import os, sys, json, collections
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200.config.presets import mt_psm_cfg
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling import rpn as R, box_head as BH, mask_head as MH, poolers as PL, detector as D
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer
from mmt_psm_b200.utils.synthetic_init import init_model
from mmt_psm_b200 import functional as Fk
from torch.profiler import ProfilerActivity, profile, record_function


def wrap(cls, name, label):
    orig = getattr(cls, name)
    def f(*a, **k):
        with record_function("REG:" + label):
            return orig(*a, **k)
    setattr(cls, name, f)

wrap(R.RPNPostProcessor, "forward", "rpn_postproc")
wrap(R.RPNLossComputation, "__call__", "rpn_loss")
wrap(BH.FastRCNNLossComputation, "subsample", "box_subsample")
wrap(BH.FastRCNNLossComputation, "__call__", "box_loss")
wrap(BH.FastRCNNLossComputation, "evaluatePSM", "psm")
wrap(BH.PostProcessor, "forward", "box_postproc")
wrap(BH.FPN2MLPFeatureExtractor, "forward", "box_feature_extractor")
wrap(BH.FPNPredictor, "forward", "box_predictor")
wrap(MH.ROIMaskHead, "forward", "mask_head")
wrap(MH.Masker, "integral_mask", "integral_mask")
wrap(D.GeneralizedRCNN, "get_fg_feature_loss", "mgd")
wrap(D.GeneralizedRCNN, "backward_deferred", "backward_deferred")

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
N = 2
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for j in range(N):
        tr.train_step(2100 + j, im, tg, vw)
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = json.load(open("/tmp/trace.json"))["traceEvents"]
regs = [(e["ts"], e["ts"] + e["dur"], e["name"][4:], e.get("tid")) for e in ev if str(e.get("name", "")).startswith("REG:")]
launch = [(e["ts"], e["name"], e.get("tid")) for e in ev if e.get("cat") == "cuda_runtime" and e["name"] in ("cudaLaunchKernel", "cudaMemcpyAsync", "cudaMemsetAsync", "cudaGraphLaunch", "cudaLaunchKernelExC")]
cnt, dur = collections.Counter(), collections.Counter()
for ts, name, tid in launch:
    inner = None
    for a, b, lab, rt in regs:
        if a <= ts < b and (inner is None or a > inner[0]):
            inner = (a, lab)
    cnt[inner[1] if inner else "(other)"] += 1
for a, b, lab, _ in regs:
    dur[lab] += b - a
print("launch-type calls per iteration by innermost region (host time of the region, ms/iter, includes nested waits):")
for lab, c in cnt.most_common():
    print("  %-24s %6.1f launches   %7.2f ms" % (lab, c / N, dur.get(lab, 0) / N / 1e3))
print("  total %.1f" % (sum(cnt.values()) / N))
