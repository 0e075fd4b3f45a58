#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15 > gpurun_out/n_pytest.log; tail -2 gpurun_out/n_pytest.log
timeout 400 python bench.py --no-cpu-baseline > gpurun_out/n_bench.json 2> gpurun_out/n_bench.log
python -c "
import json;j=json.load(open('gpurun_out/n_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), round(j['e2e']['ms_per_step'],2), j['step_ms'], j['allocator'], 'frac', round(r['frac'],3))"
