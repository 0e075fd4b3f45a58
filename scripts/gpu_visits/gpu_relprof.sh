#!/bin/bash
# Where the IR-Net iteration (config 4) spends its time + the two relation tests that changed.
mkdir -p gpurun_out
MMT_RELATION=1 timeout 400 python scripts/profile_step.py 5 gpurun_out/profile_relation_v6 > gpurun_out/profile_relation_v6.log 2>&1
echo "profile rc=$?"; grep -v Warn gpurun_out/profile_relation_v6.log | tail -14
head -32 gpurun_out/profile_relation_v6_cuda.txt | cut -c1-200
timeout 600 python -m pytest tests/test_relation_gpu.py -q --tb=short --timeout 500 -k "full_training or finite" 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_relation2_v6.log
grep -n "^E \|passed\|failed\|: in test" gpurun_out/pytest_relation2_v6.log | head -20
