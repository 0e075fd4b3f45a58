#!/bin/bash
# configs 2 (supervised only) and 4 (MT + IR-Net) on the final build
mkdir -p gpurun_out
for w in sup mt_relation; do
  timeout 120 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-roofline 2> gpurun_out/v_$w.log | tail -1 > gpurun_out/v_$w.json
  python -c "
import json; j=json.load(open('gpurun_out/v_$w.json')); print('$w value %.2f (%.2f ms) e2e %.2f (%.2f ms)' % (j['value'], j['ms_per_step'], j['e2e']['value'], j['e2e']['ms_per_step']), j['step_ms'])"
done
