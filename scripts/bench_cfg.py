def bench_cfg(start_mt=0, extra=()):
    """cfg of configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml + scripts/train_mt.sh with IR-Net off (BASELINE config 3)."""
    from mmt_psm_b200.config import cfg, TRAIN_MT_OVERRIDES
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
                       "DATALOADER.SIZE_DIVISIBILITY", 32, "SOLVER.WEIGHT_DECAY", 0.0001, "SOLVER.STEPS", (5000,), "SOLVER.MAX_ITER", 7000,
                       "MT.START_MT", start_mt] + list(extra))
    c.freeze()
    return c
