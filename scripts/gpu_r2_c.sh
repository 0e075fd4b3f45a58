#!/bin/bash
# round-2 GPU call C: ROIAlign v2 parity (suite), in-process A/B of PDL / CTA pairs / fused backward, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/c_pytest.log
tail -3 gpurun_out/c_pytest.log
timeout 600 python scripts/ab_bench.py --steps 25 > gpurun_out/c_ab.log 2>&1
cat gpurun_out/c_ab.log | grep visit
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c_bench.json 2> gpurun_out/c_bench.log
python -c "
import json;j=json.load(open('gpurun_out/c_bench.json'));r=j['roofline'];print(round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), j['step_ms'], {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items() if isinstance(v,dict)}, 'frac', round(r['frac'],3), j['clocks'])"
