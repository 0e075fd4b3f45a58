#!/bin/bash
# round-2 GPU call B: parity suite incl. CTA-pair + fused-backward tests, per-layer tile-config table, bench A/B of the knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/b_pytest.log
tail -3 gpurun_out/b_pytest.log
timeout 300 python scripts/conv_cfg_bench.py > gpurun_out/b_conv_cfg.log 2>&1
tail -3 gpurun_out/b_conv_cfg.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_bench_default.json 2> gpurun_out/b_bench_default.log
MMT_CONV_PDL=0 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_bench_nopdl.json 2> gpurun_out/b_bench_nopdl.log
MMT_CONV_CTA_GROUP=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/b_bench_cg1.json 2> gpurun_out/b_bench_cg1.log
MMT_BENCH_TRACE=1 timeout 300 python bench.py --steps 60 --no-cpu-baseline --no-roofline > gpurun_out/b_trace.json 2> gpurun_out/b_trace.log
for f in default nopdl cg1; do python -c "
import json;j=json.load(open('gpurun_out/b_bench_$f.json'));r=j['roofline'];print('$f', round(j['value'],2), round(j['ms_per_step'],2), 'e2e', round(j['e2e']['value'],2), 'conv ms', {k:round(v['ms_per_step'],2) for k,v in j['kernel_families'].items() if isinstance(v,dict)}, 'frac', round(r['frac'],3))"; done
