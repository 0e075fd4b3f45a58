#!/bin/bash
# config 4 (MT + IR-Net): per-step loss terms, shipped defaults vs the schedule knobs of the earlier build
mkdir -p gpurun_out
MMT_BENCH_LOSS_TRACE=1 timeout 60 python bench.py --workload mt_relation --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/w_a.json 2> gpurun_out/w_a.log &
wait
grep "step" gpurun_out/w_a.log | grep -c nan
MMT_BENCH_LOSS_TRACE=1 MMT_HEAD_PRIORITY=0 MMT_TEACHER_TRUNK_CTAS=148 MMT_BENCH_LOSS_DRAIN=1 timeout 60 python bench.py --workload mt_relation --steps 20 --warmup 5 --no-cpu-baseline --no-roofline > gpurun_out/w_b.json 2> gpurun_out/w_b.log
grep "step" gpurun_out/w_b.log | grep -c nan
