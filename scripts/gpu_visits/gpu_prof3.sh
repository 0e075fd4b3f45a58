#!/bin/bash
# ncu launch list of the bench command as shipped (CUDA graphs on: kernel nodes are profiled individually)
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v4.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline > gpurun_out/ncu_bench_v4.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_v4.csv; tail -c 600 gpurun_out/ncu_bench_v4.log
