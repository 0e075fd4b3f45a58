#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/ab_bench.py --steps 40 --configs base,nothread > gpurun_out/f_ab.log 2>&1
grep visit gpurun_out/f_ab.log
MMT_BENCH_TRACE=1 timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-roofline > gpurun_out/f_trace.json 2> gpurun_out/f_trace.log
awk '/^step/ {printf "%s ", $4}' gpurun_out/f_trace.log; echo
grep -E "cudaMalloc \+[1-9]" gpurun_out/f_trace.log | head -20
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.log
python -c "
import json;j=json.load(open('gpurun_out/f_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), j['step_ms'], j['allocator'], 'frac', round(r['frac'],3))"
