#!/bin/bash
# First GPU pass: SIMT-op parity, then the tcgen05 conv kernels op by op (separate processes so a
# trapped kernel cannot poison the following checks).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -k "not tc and not full_size" --timeout 600 > gpurun_out/pytest_simt.log 2>&1
echo "simt rc=$?" >> gpurun_out/pytest_simt.log
for op in fprop dgrad wgrad; do
  timeout 240 python scripts/tc_debug.py $op > gpurun_out/tc_$op.log 2>&1
  echo "rc=$?" >> gpurun_out/tc_$op.log
done
timeout 600 python -m pytest tests -m gpu -q -k "tc or full_size" --timeout 300 -s > gpurun_out/pytest_tc.log 2>&1
echo "tc rc=$?" >> gpurun_out/pytest_tc.log
tail -5 gpurun_out/pytest_simt.log; tail -15 gpurun_out/tc_fprop.log; tail -12 gpurun_out/tc_dgrad.log; tail -12 gpurun_out/tc_wgrad.log; tail -5 gpurun_out/pytest_tc.log
timeout 300 python scripts/conv_bench.py > gpurun_out/conv_bench.log 2>&1
cat gpurun_out/conv_bench.log
grep -E "passed|failed" gpurun_out/pytest_simt.log gpurun_out/pytest_tc.log
