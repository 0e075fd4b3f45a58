#!/bin/bash
# IR-Net (BASELINE config 4) visit: the relation parity tests, then a short bench of the config-4 iteration.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_relation_gpu.py -q --tb=short --timeout 600 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_relation.log
grep -n "^E \|passed\|failed\|Error\|^tests.*[0-9]: in \|^[a-z_/]*\.py:[0-9]*: in" gpurun_out/pytest_relation.log | head -60
timeout 600 python bench.py --workload mt_relation --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/bench_relation.json 2> gpurun_out/bench_relation.err
echo "bench rc=$?"; cut -c1-900 gpurun_out/bench_relation.json; tail -5 gpurun_out/bench_relation.err
