#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -f -o gpurun_out/ncu_epi_nores \
   python scripts/conv_one.py 2 200 200 64 256 1 1 0 nores fprop 2 > gpurun_out/ncu_epi_nores.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -c 2 -f -o gpurun_out/ncu_epi_res \
   python scripts/conv_one.py 2 200 200 64 256 1 1 0 res fprop 2 > gpurun_out/ncu_epi_res.log 2>&1
tail -2 gpurun_out/ncu_epi_nores.log gpurun_out/ncu_epi_res.log
ls -la gpurun_out/*.ncu-rep
