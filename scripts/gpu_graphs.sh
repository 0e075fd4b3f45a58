#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -x -q --timeout 600 2>&1 | tail -15
timeout 600 python scripts/profile_step.py 3 gpurun_out/profile_step_graphs > gpurun_out/profile_step_graphs.log 2>&1
grep -v Warn gpurun_out/profile_step_graphs.log | tail -14
