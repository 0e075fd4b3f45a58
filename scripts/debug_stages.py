import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CUDA_LAUNCH_BLOCKING"] = "1"
import mmt_psm_b200
from mmt_psm_b200.config import cfg, TRAIN_MT_OVERRIDES
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.utils.synthetic_init import init_model
exec(open(os.path.join(os.path.dirname(__file__), "bench_cfg.py")).read())
c = bench_cfg(start_mt=0)
dev = torch.device("cuda:0")
m = build_detection_model(c, is_student=True).to(dev)
init_model(m, 0)
m.eval()
size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
images, targets, views = make_batch(1234, 2, size, 3, 6, 6)
def st(name, t):
    t = t.float()
    print("%-24s shape %-22s finite %s  absmax %.4g  mean %.4g" % (name, tuple(t.shape), bool(torch.isfinite(t).all()), float(t.abs().max()), float(t.mean())), flush=True)
from mmt_psm_b200 import kernels as K
with torch.no_grad():
    x = K.nchw_to_nhwc(images.to(dev)); st("input", x)
    body = m.backbone.body
    y = body.stem.conv1(x, bn=body.stem.bn1, relu=True); st("stem conv", y)
    y = K.maxpool3x3s2(y); st("stem pool", y)
    feats = []
    for name in body.stages:
        for i, blk in enumerate(getattr(body, name)):
            y = blk(y)
            st("%s.%d" % (name, i), y)
        feats.append(y)
    ps = m.backbone.fpn(feats)
    for i, p in enumerate(ps): st("P%d" % (i + 2), p)
    lo, re = m.rpn.head(ps)
    for i, (a, b) in enumerate(zip(lo, re)): st("rpn obj %d" % i, a); st("rpn reg %d" % i, b)
print("STAGES OK")
from mmt_psm_b200.structures import to_image_list
with torch.no_grad():
    il = to_image_list(images.to(dev))
    props, _ = m.rpn(il, ps)
    print("proposals/img", [len(p) for p in props])
    fe = m.box_heads.box.feature_extractor
    pooled = fe.pooler(ps, props); st("pooled", pooled)
    x6 = fe.fc6(pooled.reshape(pooled.shape[0], -1), relu=True); st("fc6", x6)
    x7 = fe.fc7(x6, relu=True); st("fc7", x7)
    print("fc7 mean sq-norm", float((x7 ** 2).sum(1).mean()))
    cl, br = m.box_heads.box.predictor(x7); st("cls logits", cl); st("box reg", br)
    res = m.box_heads.box.post_processor((cl, br), props)
    print("detections/img", [len(r) for r in res])
