#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/i_pytest.log
tail -4 gpurun_out/i_pytest.log
timeout 600 python scripts/ab_bench.py --steps 30 --configs base,nolevels > gpurun_out/i_ab.log 2>&1
grep visit gpurun_out/i_ab.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/i_bench.json 2> gpurun_out/i_bench.log
python -c "
import json;j=json.load(open('gpurun_out/i_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), round(j['e2e']['ms_per_step'],2), j['step_ms'], j['allocator'], 'frac', round(r['frac'],3), {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items() if isinstance(v,dict)})"
