#!/bin/bash
# Mean-teacher training with the option set of the reference's scripts/train_mt.sh (kept as data in
# mmt-psm_b200/config/__init__.py: TRAIN_MT_OVERRIDES), on this package's launcher: IR-Net modules on, as in the
# reference; the one stated difference (tools/train_mean_teacher.py) is the data — the synthetic crop stream.
# Extra KEY VAL pairs on the command line override the preset, e.g.  scripts/train_mt.sh SOLVER.MAX_ITER 200
exec python tools/train_mean_teacher.py --config-file configs/pap/e2e_mask_rcnn_R_50_FPN_1x.yaml --gpuid 0 \
    --preset train_mt --synthetic "$@"
