#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -30 > gpurun_out/j_pytest.log
grep -E "passed|failed|s call" gpurun_out/j_pytest.log | head -12
timeout 400 python bench.py > gpurun_out/j_bench.json 2> gpurun_out/j_bench.log
python -c "
import json;j=json.load(open('gpurun_out/j_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), round(j['e2e']['ms_per_step'],2), j['step_ms'], j['allocator'], 'frac', round(r['frac'],3), {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items() if isinstance(v,dict)}, j['cpu_baseline'])"
MMT_BENCH_TRACE=1 timeout 300 python bench.py --steps 40 --no-cpu-baseline --no-roofline > gpurun_out/j_trace.json 2> gpurun_out/j_trace.log
awk '/^step/ {printf "%s ", $4}' gpurun_out/j_trace.log; echo; grep -E "cudaMalloc \+[1-9]" gpurun_out/j_trace.log | head
