#!/bin/bash
# ncu evidence for the bench command: (1) launch list (durations) of a short eager run, (2) DRAM traffic of every
# tcgen05 conv launch of one iteration, (3) one full-set capture of a few launches of the dominant kernel.
mkdir -p gpurun_out
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v3.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-graphs > gpurun_out/ncu_bench_v3.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches_v3.csv
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
   --clock-control none -k regex:"conv_tc_kernel|wgrad_tc_kernel" -s 500 -c 500 --csv --log-file gpurun_out/conv_traffic_v3.csv \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-graphs > gpurun_out/ncu_traffic_v3.log 2>&1
echo "ncu traffic rc=$?"; wc -l gpurun_out/conv_traffic_v3.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 520 -c 8 -f -o gpurun_out/prof_conv_v3 \
   python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-roofline --no-graphs > gpurun_out/ncu_full_v3.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/ | tail -8
