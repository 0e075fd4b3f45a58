#!/bin/bash
# config 4 through the launcher (train -> checkpoint -> resume) and a longer-warm-up bench of the config-4 iteration
mkdir -p gpurun_out
bash scripts/gpu_launcher.sh > gpurun_out/launcher_relation_v6.log 2>&1
tail -14 gpurun_out/launcher_relation_v6.log | cut -c1-260
MMT_BENCH_TRACE=1 timeout 600 python bench.py --workload mt_relation --no-cpu-baseline --no-roofline --warmup 5 --steps 60 > gpurun_out/bench_relation_trace_v6.json 2> gpurun_out/bench_relation_trace_v6.err
grep "^step dev" gpurun_out/bench_relation_trace_v6.err | awk '{printf "%s ", $4} END {print ""}'
timeout 600 python bench.py --workload mt_relation --no-cpu-baseline --no-roofline --warmup 40 --steps 40 > gpurun_out/bench_relation_w40_v6.json 2> gpurun_out/bench_relation_w40_v6.err
echo "w40 rc=$?"; cut -c1-330 gpurun_out/bench_relation_w40_v6.json; python -c "
import json; d=json.load(open('gpurun_out/bench_relation_w40_v6.json')); print('value', d['value'], d['ms_per_step'], 'e2e', d['e2e'])"
