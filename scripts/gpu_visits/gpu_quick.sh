#!/bin/bash
# Quick GPU visit while iterating: parity suite, free-running step time, 30-iteration loss trace.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_gpu.log
grep -n "Error\|^E \|passed\|failed" gpurun_out/pytest_gpu.log | head -30
MMT_PROFILE_QUICK=1 timeout 600 python scripts/profile_step.py 8 2>&1 | grep "free-running"
timeout 300 python scripts/loss_trace.py 30 2000 2>/dev/null | cut -c1-230 | tail -2
