#!/bin/bash
# Round-end visit after the IR-Net work: step-level + relation parity suites, smoke, bench (config 3 headline, config 4, reference arm).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests/test_relation_gpu.py tests/test_model_gpu.py -q --tb=short --timeout 900 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_gpu_v6.log
grep -n "^E \|passed\|failed\|: in test" gpurun_out/pytest_gpu_v6.log | head -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v6.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke_v6.log
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_v6.json
timeout 600 python bench.py --workload mt_relation --no-cpu-baseline > gpurun_out/bench_relation_v6.json 2> gpurun_out/bench_relation_v6.err
echo "relation rc=$?"; cut -c1-900 gpurun_out/bench_relation_v6.json; tail -3 gpurun_out/bench_relation_v6.err
timeout 600 python bench.py --impl reference --workload mt_relation --steps 1 --warmup 0 > gpurun_out/bench_relation_ref_v6.json 2> gpurun_out/bench_relation_ref_v6.err
echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_relation_ref_v6.json
