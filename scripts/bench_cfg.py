def bench_cfg(start_mt=0, extra=()):
    """cfg of configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml + scripts/train_mt.sh with IR-Net off (BASELINE config 3)."""
    from mmt_psm_b200.config.presets import mt_psm_cfg
    return mt_psm_cfg(start_mt, extra=extra)
