"""The configuration the headline metric is quoted on, without needing the reference's yaml on disk:
`configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml` (reference, keys read by the path) merged with the
`scripts/train_mt.sh` overrides (reference scripts/train_mt.sh:4-20), IR-Net off (BASELINE config 3)."""
from . import TRAIN_MT_OVERRIDES, cfg

# the values of reference configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml that differ from config/defaults.py
PAP_R50_FPN = [
    "MODEL.META_ARCHITECTURE", "GeneralizedRCNN", "MODEL.MASK_ON", True,
    "MODEL.BACKBONE.CONV_BODY", "R-50-FPN", "MODEL.BACKBONE.OUT_CHANNELS", 256,
    "MODEL.RPN.USE_FPN", True, "MODEL.RPN.ANCHOR_STRIDE", (4, 8, 16, 32, 64),
    "MODEL.RPN.PRE_NMS_TOP_N_TRAIN", 2000, "MODEL.RPN.PRE_NMS_TOP_N_TEST", 1000,
    "MODEL.RPN.POST_NMS_TOP_N_TEST", 1000, "MODEL.RPN.FPN_POST_NMS_TOP_N_TEST", 1000,
    "MODEL.ROI_HEADS.USE_FPN", True,
    "MODEL.ROI_BOX_HEAD.POOLER_RESOLUTION", 7, "MODEL.ROI_BOX_HEAD.POOLER_SCALES", (0.25, 0.125, 0.0625, 0.03125),
    "MODEL.ROI_BOX_HEAD.POOLER_SAMPLING_RATIO", 2, "MODEL.ROI_BOX_HEAD.FEATURE_EXTRACTOR", "FPN2MLPFeatureExtractor",
    "MODEL.ROI_BOX_HEAD.PREDICTOR", "FPNPredictor",
    "MODEL.ROI_MASK_HEAD.POOLER_SCALES", (0.25, 0.125, 0.0625, 0.03125),
    "MODEL.ROI_MASK_HEAD.FEATURE_EXTRACTOR", "MaskRCNNFPNFeatureExtractor",
    "MODEL.ROI_MASK_HEAD.PREDICTOR", "MaskRCNNC4Predictor", "MODEL.ROI_MASK_HEAD.POOLER_RESOLUTION", 14,
    "MODEL.ROI_MASK_HEAD.POOLER_SAMPLING_RATIO", 2, "MODEL.ROI_MASK_HEAD.RESOLUTION", 28,
    "MODEL.ROI_MASK_HEAD.SHARE_BOX_FEATURE_EXTRACTOR", False,
    "DATALOADER.SIZE_DIVISIBILITY", 32,
    "SOLVER.WEIGHT_DECAY", 0.0001, "SOLVER.STEPS", (5000,), "SOLVER.MAX_ITER", 7000,
]


def mt_psm_cfg(start_mt=1000, relation=False, extra=()):
    """Frozen cfg of `sh scripts/train_mt.sh`; `relation=False` switches the IR-Net modules off (BASELINE config 3)."""
    c = cfg.clone()
    c.merge_from_list(PAP_R50_FPN)
    c.merge_from_list(TRAIN_MT_OVERRIDES)
    c.merge_from_list(["MODEL.RELATION_NMS.USE_RELATION_NMS", relation, "MODEL.RELATION_MASK.USE_RELATION", relation,
                       "MT.START_MT", start_mt] + list(extra))
    c.freeze()
    return c


def supervised_cfg(extra=()):
    """BASELINE configs 1/2: supervised-only Mask R-CNN R50-FPN (MT.LAMBDA = 0, relation off)."""
    return mt_psm_cfg(extra=["MT.LAMBDA", 0.0, "DATASETS.NO_LABEL", False] + list(extra))
