#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_gpu.log
grep -n "Error\|^E \|passed\|failed" gpurun_out/pytest_gpu.log | head -30
timeout 600 python scripts/cprofile_step.py > gpurun_out/cprofile_step.log 2>&1
