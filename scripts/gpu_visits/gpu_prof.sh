#!/bin/bash
# ncu launch list of a short bench run + one full-set capture of the dominant conv kernel + loss trace.
mkdir -p gpurun_out
timeout 600 python scripts/loss_trace.py 16 2000 > gpurun_out/loss_trace_lr_full.log 2>&1
timeout 600 python scripts/loss_trace.py 16 0 > gpurun_out/loss_trace_lr_warmup.log 2>&1
tail -4 gpurun_out/loss_trace_lr_full.log; tail -2 gpurun_out/loss_trace_lr_warmup.log
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 300 -c 6 -f -o gpurun_out/prof_conv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/
