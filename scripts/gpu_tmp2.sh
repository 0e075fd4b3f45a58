#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/launch_census.py > gpurun_out/launch_census.log 2>&1
grep -v Warn gpurun_out/launch_census.log | tail -17
