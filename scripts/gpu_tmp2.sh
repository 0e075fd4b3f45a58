#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/profile_step.py 5 gpurun_out/profile_step_graphs > gpurun_out/profile_step_graphs.log 2>&1
grep -v Warn gpurun_out/profile_step_graphs.log | tail -4
