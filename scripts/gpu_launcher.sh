#!/bin/bash
# launcher smoke: 12 iterations (MT from iteration 3), checkpoint every 5, then resume to 16
rm -rf /tmp/run_demo
timeout 600 bash scripts/train_mt.sh SOLVER.MAX_ITER 12 MT.START_MT 2 SOLVER.CHECKPOINT_PERIOD 5 INPUT.MIN_SIZE_TRAIN 320 INPUT.MAX_SIZE_TRAIN 320 OUTPUT_DIR /tmp/run_demo 2>&1 | grep -v Warning | tail -6
ls /tmp/run_demo /tmp/run_demo/teacher
timeout 600 bash scripts/train_mt.sh SOLVER.MAX_ITER 16 MT.START_MT 2 SOLVER.CHECKPOINT_PERIOD 5 INPUT.MIN_SIZE_TRAIN 320 INPUT.MAX_SIZE_TRAIN 320 OUTPUT_DIR /tmp/run_demo 2>&1 | grep -v Warning | grep -i "loading\|training\|iter:\|Total" | tail -6
