"""The configuration the headline metric is quoted on, without needing the reference's yaml on disk:
`configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml` (reference, keys read by the path) merged with the
`scripts/train_mt.sh` overrides (reference scripts/train_mt.sh:4-20); `relation=` switches the IR-Net modules
(BASELINE config 3 = off, config 4 = on, which is what the reference's yaml itself selects)."""
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

# the IR-Net blocks of the same yaml (reference configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml:35-85)
PAP_RELATION = [
    "MODEL.RELATION_NMS.IOU_METHOD", "n", "MODEL.RELATION_NMS.USE_IOU", False, "MODEL.RELATION_NMS.MUTRELATION", False,
    "MODEL.RELATION_NMS.FIRST_N", 90, "MODEL.RELATION_NMS.THREAD", (0.1,), "MODEL.RELATION_NMS.ROI_FEAT_DIM", 1024,
    "MODEL.RELATION_NMS.APPEARANCE_FEAT_DIM", 128, "MODEL.RELATION_NMS.GEO_FEAT_DIM", 64,
    "MODEL.RELATION_NMS.FC_DIM", (64, 16), "MODEL.RELATION_NMS.GROUP", 16, "MODEL.RELATION_NMS.HID_DIM", (1024, 1024, 128),
    "MODEL.RELATION_NMS.CLASS_AGNOSTIC", False, "MODEL.RELATION_NMS.FG_THREAD", 0.1, "MODEL.RELATION_NMS.MERGE_METHOD", 0,
    "MODEL.RELATION_NMS.POS_NMS", 0.55, "MODEL.RELATION_NMS.CLS_WISE_RELATION", False, "MODEL.RELATION_NMS.TOPK", 40,
    "MODEL.RELATION_NMS.TAG", "CIAM", "MODEL.RELATION_NMS.APPEARANCE_INTER", True, "MODEL.RELATION_NMS.CONCAT", False,
    "MODEL.RELATION_MASK.BINARY", False, "MODEL.RELATION_MASK.USE_PRE_FEATURE", False, "MODEL.RELATION_MASK.PRE_NORM", False,
    "MODEL.RELATION_MASK.DEEP_SUPER", True, "MODEL.RELATION_MASK.TYPE", "CIAM", "MODEL.RELATION_MASK.SAME_PREDICTOR", False,
    "MODEL.RELATION_MASK.CAM", False, "MODEL.RELATION_MASK.CIAM", True, "MODEL.RELATION_MASK.TRAIN_CENTER_ONLY", False,
    "MODEL.RELATION_MASK.PROTO", False, "MODEL.RELATION_MASK.ALPHA", 0.5, "MODEL.RELATION_MASK.RANK", True,
    "MODEL.RELATION_MASK.XY_COOR", True, "MODEL.RELATION_MASK.NORM", -1, "MODEL.RELATION_MASK.IOU_COOR", False,
    "MODEL.RELATION_MASK.FEATURE_EXTRACTOR", "RoiAlignMaskFeatureExtractor", "MODEL.RELATION_MASK.EXTRACTOR_CHANNEL", 16,
    "MODEL.RELATION_MASK.APPEARANCE_FEAT_DIM", 128, "MODEL.RELATION_MASK.GEO_FEAT_DIM", 64,
    "MODEL.RELATION_MASK.FC_DIM", (64, 16), "MODEL.RELATION_MASK.GROUP", 16, "MODEL.RELATION_MASK.HID_DIM", (1024, 1024),
    "MODEL.RELATION_MASK.TOPK", 128, "MODEL.RELATION_MASK.CLSWIZE", True,
]


def mt_psm_cfg(start_mt=1000, relation=False, extra=()):
    """Frozen cfg of `sh scripts/train_mt.sh`; `relation=False` switches the IR-Net modules off (BASELINE config 3)."""
    c = cfg.clone()
    c.merge_from_list(PAP_R50_FPN)
    c.merge_from_list(PAP_RELATION)
    c.merge_from_list(TRAIN_MT_OVERRIDES)
    c.merge_from_list(["MODEL.RELATION_NMS.USE_RELATION_NMS", relation, "MODEL.RELATION_MASK.USE_RELATION", relation,
                       "MT.START_MT", start_mt] + list(extra))
    c.freeze()
    return c


def supervised_cfg(extra=()):
    """BASELINE configs 1/2: supervised-only Mask R-CNN R50-FPN (MT.LAMBDA = 0, relation off)."""
    return mt_psm_cfg(extra=["MT.LAMBDA", 0.0, "DATASETS.NO_LABEL", False] + list(extra))
