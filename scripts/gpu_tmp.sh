#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -f -o gpurun_out/ncu_epi_res2 \
   python scripts/conv_one.py 2 200 200 64 256 1 1 0 res fprop 2 > gpurun_out/ncu_epi_res2.log 2>&1
tail -n 2 gpurun_out/ncu_epi_res2.log
timeout 600 python scripts/conv_bench.py --all > gpurun_out/conv_bench.log 2>&1
tail -n 1 gpurun_out/conv_bench.log
