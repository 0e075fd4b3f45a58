#!/bin/bash
# Round-end style visit: GPU parity suites (ops, step, IR-Net), smoke, bench (ours + reference arm), secondary workloads incl. config 4.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/nproc.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | grep -v "^E   *\[" > gpurun_out/pytest_gpu.log
grep -n "^E \|passed\|failed" gpurun_out/pytest_gpu.log | head -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 600 python bench.py --workload sup --no-cpu-baseline > gpurun_out/bench_sup.json 2> gpurun_out/bench_sup.err
echo "sup rc=$?"; cut -c1-330 gpurun_out/bench_sup.json
timeout 600 python bench.py --size 1024 --no-cpu-baseline > gpurun_out/bench_1024.json 2> gpurun_out/bench_1024.err
echo "1024 rc=$?"; cut -c1-330 gpurun_out/bench_1024.json
timeout 600 python bench.py --workload mt_relation --no-cpu-baseline --warmup 20 --steps 40 > gpurun_out/bench_relation.json 2> gpurun_out/bench_relation.err
echo "relation rc=$?"; cut -c1-330 gpurun_out/bench_relation.json
