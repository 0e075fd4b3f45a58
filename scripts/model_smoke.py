"""GPU smoke of the full path: supervised step, then a mean-teacher step (teacher + student-unlabeled + fused SGD/EMA)."""
import os
import sys
import time

import logging
import torch
logging.basicConfig(level=logging.INFO)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mmt_psm_b200  # noqa
from mmt_psm_b200 import lib
from mmt_psm_b200.config import cfg, TRAIN_MT_OVERRIDES
from mmt_psm_b200.data.synthetic import make_batch
from mmt_psm_b200.engine import MTtrainer
from mmt_psm_b200.modeling.detector import build_detection_model
from mmt_psm_b200.solver import make_lr_scheduler, make_optimizer

size = int(sys.argv[1]) if len(sys.argv) > 1 else 256
c = cfg.clone()
c.merge_from_list(TRAIN_MT_OVERRIDES)
c.merge_from_list(["MODEL.RPN.USE_FPN", True, "MODEL.ROI_HEADS.USE_FPN", True, "MODEL.BACKBONE.CONV_BODY", "R-50-FPN",
                   "MODEL.BACKBONE.OUT_CHANNELS", 256, "MODEL.RPN.ANCHOR_STRIDE", (4, 8, 16, 32, 64),
                   "MODEL.RPN.PRE_NMS_TOP_N_TRAIN", 2000, "MODEL.RPN.PRE_NMS_TOP_N_TEST", 1000,
                   "MODEL.RPN.POST_NMS_TOP_N_TEST", 1000, "MODEL.RPN.FPN_POST_NMS_TOP_N_TEST", 1000,
                   "MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", 7, "MODEL.ROI_BOX_HEAD.POOLER_SCALES", (0.25, 0.125, 0.0625, 0.03125),
                   "MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO", 2, "MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR", "FPN2MLPFeatureExtractor",
                   "MODEL.ROI_BOX_HEAD.PREDICTOR", "FPNPredictor", "MODEL.ROI_MASK_HEAD.POOLER_SCALES", (0.25, 0.125, 0.0625, 0.03125),
                   "MODEL.ROI_MASK_HEAD.FEATURE_EXTRACTOR", "MaskRCNNFPNFeatureExtractor", "MODEL.ROI_MASK_HEAD.PREDICTOR", "MaskRCNNC4Predictor",
                   "MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION", 14, "MODEL.ROI_MASK_HEAD.POOLER_SAMPLING_RATIO", 2,
                   "MODEL.ROI_MASK_HEAD.RESOLUTION", 28, "MODEL.ROI_MASK_HEAD.SHARE_BOX_FEATURE_EXTRACTOR", False,
                   "MODEL.MASK_ON", True, "MODEL.RELATION_NMS.USE_RELATION_NMS", False, "MODEL.RELATION_MASK.USE_RELATION", False,
                   "DATALOADER.SIZE_DIVISIBILITY", 32, "MT.START_MT", 0])
c.freeze()
dev = torch.device("cuda:0")
torch.manual_seed(0)
student = build_detection_model(c, is_student=True).to(dev)
teacher = build_detection_model(c, is_teacher=True).to(dev)
from mmt_psm_b200.utils.synthetic_init import init_model
init_model(student, 0)
teacher.load_state_dict(student.state_dict())
opt = make_optimizer(c, student, teacher)
sched = make_lr_scheduler(c, opt)
trainer = MTtrainer(student, teacher, None, opt, sched, None, None, 1000, c)
student.train()
teacher.eval()
images, targets, views = make_batch(1234, 2, size, 3, n_cyto=6 if size < 600 else 20, n_nuc=6 if size < 600 else 20)
lib.reset_launch_count()
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 4):
    torch.cuda.synchronize()
    t0 = time.time()
    ld, total = trainer.train_step(it, images, targets, views if it > 0 else None)
    torch.cuda.synchronize()
    print("iter", it, "%.1f ms" % ((time.time() - t0) * 1e3), {k: round(float(v), 5) for k, v in ld.items()},
          "launches", lib.launch_count(), flush=True)
    assert all(torch.isfinite(v).all() for v in ld.values())
print("max mem MB", torch.cuda.max_memory_allocated() / 2**20)
print("SMOKE OK")
