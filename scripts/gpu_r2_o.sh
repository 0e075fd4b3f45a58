#!/bin/bash
for v in 1 2 0 1 2; do
MMT_BENCH_STAGE=$v timeout 300 python bench.py --no-cpu-baseline --no-roofline --steps 30 > gpurun_out/o_$v.json 2>/dev/null
python -c "
import json;j=json.load(open('gpurun_out/o_$v.json'));print('stage=$v dev', round(j['ms_per_step'],2), 'e2e', round(j['e2e']['ms_per_step'],2), j['step_ms']['e2e'])"
done
