#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/e_pytest.log
tail -3 gpurun_out/e_pytest.log
timeout 600 python scripts/ab_bench.py --steps 25 --configs base,nothread > gpurun_out/e_ab.log 2>&1
grep visit gpurun_out/e_ab.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/e_bench.json 2> gpurun_out/e_bench.log
python -c "
import json;j=json.load(open('gpurun_out/e_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), j['step_ms'], 'frac', round(r['frac'],3))"
timeout 300 python scripts/gap_trace.py > gpurun_out/e_gap_trace.log 2>&1; grep -E "window|iterations|total idle" gpurun_out/e_gap_trace.log
