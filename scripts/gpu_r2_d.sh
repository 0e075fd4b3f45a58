#!/bin/bash
mkdir -p gpurun_out
timeout 400 python scripts/gap_trace.py > gpurun_out/d_gap_trace.log 2>&1
timeout 400 python scripts/profile_step.py 3 gpurun_out/d_profile_step > gpurun_out/d_profile_step.log 2>&1
tail -30 gpurun_out/d_gap_trace.log; cat gpurun_out/d_profile_step.log | tail -15
