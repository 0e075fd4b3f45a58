#!/bin/bash
# One GPU visit: parity tests, smoke, bench (both arms optional), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -6 gpurun_out/smoke.log
timeout 900 python bench.py --steps ${STEPS:-10} --warmup ${WARMUP:-3} ${BENCH_ARGS} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"
tail -3 gpurun_out/bench.err
cat gpurun_out/bench.json
